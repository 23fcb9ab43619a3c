#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export HDRVAE_FUSE_GN=1 HDRVAE_NO_GRAPH=1
python tools/profile_decode.py 4 128 > gpurun_out/fused_plain.log 2>&1 || { tail -5 gpurun_out/fused_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/prof_conv_fused -f python tools/profile_decode.py 4 128 > gpurun_out/ncu_fused.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_conv_fused.ncu-rep
