#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( HDRVAE_X16=1 timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/x16_pytest_all.log 2>&1
echo "pytest (x16, all) rc=$?"; tail -12 gpurun_out/x16_pytest_all.log | cut -c1-300
( HDRVAE_X16=0 timeout 1200 python -m pytest tests/test_gpu_decode.py tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/x16_pytest_off.log 2>&1
echo "pytest (x16 off, decode+multi) rc=$?"; tail -3 gpurun_out/x16_pytest_off.log | cut -c1-300
