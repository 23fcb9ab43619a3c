#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 120 -k "attention" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -x --timeout 120 2>&1 | tail -3
timeout 900 python tools/big_decode.py 256 --check 2>&1 | tail -4
timeout 1200 python tools/big_decode.py 512 2>&1 | tail -2
grep -E "attention|TOTAL" gpurun_out/profile_big_512.tsv
