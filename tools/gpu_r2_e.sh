#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "groupnorm" ) > gpurun_out/r2e_gn_tests.log 2>&1
echo "gn tests rc=$?"; tail -3 gpurun_out/r2e_gn_tests.log
python tools/profile_decode.py 4 128 gpurun_out/r2e_per_op_c2.tsv > gpurun_out/r2e_profile.log 2>&1; tail -1 gpurun_out/r2e_profile.log
grep groupnorm gpurun_out/r2e_per_op_c2.tsv | awk -F'\t' '{s+=$2} END {print "groupnorm total ms (burst clocks):", s}'
grep groupnorm gpurun_out/r2d_per_op_c2.tsv 2>/dev/null | awk -F'\t' '{s+=$2} END {print "previous:", s}'
bash tools/gpu_r2_ncu.sh
