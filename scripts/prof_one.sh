#!/bin/bash
# ncu --set full + source page of ONE kernel: bash scripts/prof_one.sh <tag> <kernel regex> <prof_one.py args...>
TAG=$1; KREG=$2; shift 2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$KREG" --launch-skip 1 --launch-count 1 -o gpurun_out/${TAG} -f python scripts/prof_one.py "$@" > gpurun_out/${TAG}.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_src.csv
rm -f gpurun_out/${TAG}.ncu-rep
