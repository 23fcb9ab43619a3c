#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_upscaler.py -q -m gpu -x --timeout 600 > gpurun_out/up_test.log 2>&1; echo "upscaler test exit $?"; tail -30 gpurun_out/up_test.log
