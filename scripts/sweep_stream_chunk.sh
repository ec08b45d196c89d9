for sc in 0 256 64; do
  python bench.py --no-cpu-baseline --no-eager-baseline --stream-chunk $sc > gpurun_out/sc.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/sc.json").read().strip().splitlines()[-1])
print("stream_chunk $sc: series/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "e2e ms", round(d["e2e"]["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"])
PY
done
