#!/bin/bash
# quick A/B sweeps on the final kernels: dual chain on/off, micro-batch
for cfg in "64 1" "64 0" "128 1" "96 1" "48 1" "32 1"; do
  set -- $cfg
  echo "micro_batch=$1 dual_chain=$2"
  SVB_DUAL_CHAIN=$2 python bench.py --no-cpu-baseline --no-eager-baseline --micro-batch $1 > gpurun_out/sw.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/sw.json").read().strip().splitlines()[-1])
print("   series/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"]["sm_mhz"], {k: round(v,2) for k,v in d["kernel_ms_per_step"].items() if k in ("dwconv_ln","gemm")})
PY
done
