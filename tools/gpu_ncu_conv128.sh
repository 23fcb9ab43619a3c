#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python tools/one_conv.py 1024 128 128 3 1 0 4 > gpurun_out/one_conv_plain.log 2>&1 || { tail -5 gpurun_out/one_conv_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_conv128_b4 -f python tools/one_conv.py 1024 128 128 3 1 0 4 > gpurun_out/ncu_conv128.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_conv128_b4.ncu-rep
