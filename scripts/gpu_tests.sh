#!/bin/bash
# Run each GPU test file in its own process (a device-side trap must not poison the next file).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in ${@:-tests/test_gpu_abi_errors.py tests/test_gpu_k0.py tests/test_gpu_k1.py tests/test_gpu_k3.py tests/test_gpu_k4.py tests/test_gpu_dataset.py tests/test_gpu_localization.py tests/test_gpu_gemm.py tests/test_gpu_model.py}; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout=600 -x > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $f exit $r"; tail -n 25 gpurun_out/$name.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
