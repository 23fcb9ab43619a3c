#!/bin/bash
# quick validation + C2 timing after a kernel change
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for f in test_gpu_kernels test_gpu_decode test_gpu_upscaler; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; tail -3 gpurun_out/$f.log
done
timeout 300 python tools/profile_decode.py 4 128 gpurun_out/profile_c2.tsv > gpurun_out/profile_c2.log 2>&1; echo "profile exit $?"; tail -1 gpurun_out/profile_c2.log
timeout 300 python tools/up_bench.py 1024 1024 23 > gpurun_out/up_bench.log 2>&1; tail -1 gpurun_out/up_bench.log
