#!/bin/bash
# Final verification in ONE call on a fresh box: every GPU test in one process (what the driver runs), smoke(), both bench arms.
#   bash scripts/final_check.sh <tag>     -> gpurun_out/<tag>_pytest_gpu.log, <tag>_smoke.log, <tag>_bench.json, <tag>_bench_reference_arm.json
TAG=${1:-final}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -s -m gpu -p no:cacheprovider --durations=12 ) > gpurun_out/${TAG}_pytest_gpu_full.log 2>&1; echo "pytest rc=$?"
grep -E "^\[coords\]|^\[crops\]|passed|failed|error" gpurun_out/${TAG}_pytest_gpu_full.log | tail -n 40 > gpurun_out/${TAG}_pytest_gpu.log
tail -n 20 gpurun_out/${TAG}_pytest_gpu_full.log >> gpurun_out/${TAG}_pytest_gpu.log
rm -f gpurun_out/${TAG}_pytest_gpu_full.log
tail -n 45 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/${TAG}_bench.json
