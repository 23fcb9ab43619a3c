#!/bin/bash
# ncu --set full of two fused-GroupNorm conv launches of a C2 decode (256 -> 256 @ 4 x 512^2: conv1, then conv2 with the
# 16-bit in-place residual), graphs off; after the same command ran without ncu
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export HDRVAE_NO_GRAPH=1
P="python tools/profile_decode.py 4 128"
$P > gpurun_out/xf_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 138 -c 2 -o gpurun_out/r2_prof_conv_xf -f $P > gpurun_out/xf_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r2_prof_conv_xf.ncu-rep
