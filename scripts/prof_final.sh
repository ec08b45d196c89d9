mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r01l.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r01l.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k1_|k3_crop|k4_" -o gpurun_out/prof_k1k3k4_r01l -f python scripts/prof_k1k3.py 256 0 > gpurun_out/ncu_k1k3_r01l.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/prof_k1k3k4_r01l.ncu-rep gpurun_out/launches_r01l.csv
