#!/bin/bash
# ncu --set full of one phase launch of an upsample conv (256->256 @4x512^2 -> 1024^2, GroupNorm statistics)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
A="python tools/one_conv.py 512 256 256 3 0 1 4"
$A > gpurun_out/r2_ncu_plain_up.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 8 -c 1 -o gpurun_out/r2_prof_conv_up256 -f $A > gpurun_out/r2_ncu_up.log 2>&1
echo "ncu up exit $?"
ls -la gpurun_out/r2_prof_conv_up256.ncu-rep
