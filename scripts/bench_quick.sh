#!/bin/bash
# short bench without the CPU / eager legs + headline summary: bash scripts/bench_quick.sh <tag> [extra bench.py args]
TAG=$1; shift
python bench.py --no-cpu-baseline --no-eager-baseline "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("series/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["clocks"], "launches", d["gpu_launches"])
print({k: round(v,2) for k,v in d["kernel_ms_per_step"].items()})
PY
