#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -k "bit_identical or residual_stream" ) > gpurun_out/newtests.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/newtests.log | cut -c1-300
