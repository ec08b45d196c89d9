# Round-2 evidence set (run under gpurun, one GPU; every ncu pass runs only after the plain command has exited 0 without ncu):
#   1. plain bench (CUDA events)                         -> gpurun_out/r02_bench.json
#   2. ncu launch list of the same command               -> gpurun_out/r02_launches.csv
#   3. ncu --set full, one launch per layer kernel of the default forward at micro-batch 64 (scripts/prof_layers.py 64 0)
#   4. ncu --set full of K0 / K1 / K3 / K4 on the bench shapes
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 200 python scripts/prof_layers.py 64 5 > gpurun_out/${TAG}_layer_timings.txt 2>&1; echo "layers rc=$?"
timeout 200 python scripts/prof_k1k3.py 256 5 >> gpurun_out/${TAG}_layer_timings.txt 2>&1; echo "k1k3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gemm_kernel|mlp_fused_kernel|dwconv_rawtc_kernel|ln_stat_finalize_kernel|stem_ln_kernel" -o gpurun_out/${TAG}_prof_layers -f python scripts/prof_layers.py 64 0 > gpurun_out/${TAG}_ncu_layers.log 2>&1; echo "full layers rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k0_|k1_|k3_crop|k4_" -o gpurun_out/${TAG}_prof_k0k1k3k4 -f python scripts/prof_k1k3.py 256 0 > gpurun_out/${TAG}_ncu_k1k3.log 2>&1; echo "full k rc=$?"
# summarise on the box (the reports are too large to travel back: gpurun returns at most 64 MiB), then drop them
python scripts/summarise_launches.py gpurun_out/${TAG}_launches.csv ${TAG} "ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline" > /dev/null; echo "sum launches rc=$?"
python scripts/summarise_ncu_layers.py gpurun_out/${TAG}_prof_layers.ncu-rep ${TAG} 64; echo "sum layers rc=$?"
python scripts/summarise_ncu.py gpurun_out/${TAG}_prof_k0k1k3k4.ncu-rep profiles/${TAG}_ncu_full_k0k1k3k4.csv; echo "sum k rc=$?"
cp profiles/${TAG}_* gpurun_out/ 2>/dev/null
rm -f gpurun_out/${TAG}_prof_layers.ncu-rep gpurun_out/${TAG}_prof_k0k1k3k4.ncu-rep gpurun_out/${TAG}_launches.csv
ls -la gpurun_out/${TAG}_*
