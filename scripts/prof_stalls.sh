#!/bin/bash
# Source-level stall capture of the layer kernels (one launch each at micro-batch 64); writes SASS-level CSVs.
TAG=${1:-r02k}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"dwconv_raw_kernel|gemm_kernel" -o gpurun_out/${TAG}_stalls -f python scripts/prof_layers.py 64 0 > gpurun_out/${TAG}_stalls.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_stalls.ncu-rep --page raw --csv > gpurun_out/${TAG}_stalls_raw.csv 2>/dev/null
for id in 0 1 2 3 4 5 6 7 8 9 10 11; do
  ncu -i gpurun_out/${TAG}_stalls.ncu-rep --page source --csv --launch-skip $id --launch-count 1 > gpurun_out/${TAG}_src_$id.csv 2>/dev/null
done
ls -la gpurun_out/${TAG}_*
gzip -f gpurun_out/${TAG}_src_*.csv
rm -f gpurun_out/${TAG}_stalls.ncu-rep
