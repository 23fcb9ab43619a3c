#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_decode.py -q -m gpu -x --timeout 300 -k "row_tiled or full_decode" > gpurun_out/rows_test.log 2>&1; echo "rows test exit $?"; tail -25 gpurun_out/rows_test.log
