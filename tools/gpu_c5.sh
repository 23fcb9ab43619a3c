#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_upscaler.py -q -m gpu -x --timeout 600 -k "c5_pipeline" > gpurun_out/c5_test.log 2>&1; echo "c5 test exit $?"; tail -5 gpurun_out/c5_test.log
timeout 600 python tools/c5_pipeline.py > gpurun_out/c5_pipeline.log 2>&1; echo "c5 pipeline exit $?"; tail -3 gpurun_out/c5_pipeline.log
timeout 600 python tools/c3_batch.py > gpurun_out/c3_batch.log 2>&1; echo "c3 exit $?"; tail -5 gpurun_out/c3_batch.log
