#!/bin/bash
# ncu --set full of the in-place-residual 128-channel conv (128->128 @1024^2, B = 4, fp32 residual + GroupNorm statistics)
# and of the same conv without residual, to see what bounds the residual epilogue
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
A="python tools/one_conv.py 1024 128 128 3 1 0 4"
$A > gpurun_out/r2_ncu_plain_res.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r2_prof_conv128_res -f $A > gpurun_out/r2_ncu_res.log 2>&1
echo "ncu res exit $?"
B="python tools/one_conv.py 1024 128 128 3 0 0 4"
$B > gpurun_out/r2_ncu_plain_nores.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r2_prof_conv128_nores -f $B > gpurun_out/r2_ncu_nores.log 2>&1
echo "ncu nores exit $?"
ls -la gpurun_out/r2_prof_conv128*.ncu-rep
