#!/bin/bash
# ncu --set full of the slab conv on a decoder shape (256->256 @512^2, B=4, GroupNorm statistics): $1 = output name
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python tools/one_conv.py 512 256 256 3 0 0 4 > gpurun_out/one_conv_plain.log 2>&1 || { tail -5 gpurun_out/one_conv_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_conv256_slab -f python tools/one_conv.py 512 256 256 3 0 0 4 > gpurun_out/ncu_conv256.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_conv256_slab.ncu-rep
