#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for f in test_gpu_epilogue test_gpu_decode; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x --timeout 300 > gpurun_out/$f.log 2>&1; echo "$f exit $?"; tail -2 gpurun_out/$f.log
done
timeout 300 python tools/profile_decode.py 4 128 gpurun_out/profile_c2.tsv | tail -1; grep "epilogue" gpurun_out/profile_c2.tsv | cut -c1-100
