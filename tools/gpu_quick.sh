#!/bin/bash
# quick GPU check: all parity tests (stop at first failure), profile
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for f in test_gpu_kernels test_gpu_epilogue test_gpu_decode; do
timeout 600 python -m pytest tests/$f.py -q -m gpu -x --timeout 120 > gpurun_out/quick_$f.log 2>&1; echo "$f exit $?"; tail -3 gpurun_out/quick_$f.log
done
timeout 300 python tools/profile_decode.py 4 128 gpurun_out/profile_c2.tsv 2>&1 | tail -2
grep -E "epilogue|attention|TOTAL" gpurun_out/profile_c2.tsv
timeout 300 python tools/iter_times.py 4 128 8 2>&1 | head -3
