mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=12 ) > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 30 gpurun_out/final_pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench26_reference.json 2> gpurun_out/bench26_reference.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench26.json 2> gpurun_out/bench26.err; echo "bench rc=$?"
cat gpurun_out/bench26.json | cut -c1-600
