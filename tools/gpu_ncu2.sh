#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python tools/one_conv.py 1024 256 128 1 0 0 1 > gpurun_out/one_conv_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_nin -f python tools/one_conv.py 1024 256 128 1 0 0 1 > gpurun_out/ncu_nin.log 2>&1
echo "ncu nin exit $?"
python tools/one_conv.py 1024 128 128 3 1 0 1 > gpurun_out/one_conv_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_conv128_res_v2 -f python tools/one_conv.py 1024 128 128 3 1 0 1 > gpurun_out/ncu_conv2.log 2>&1
echo "ncu conv128 exit $?"
