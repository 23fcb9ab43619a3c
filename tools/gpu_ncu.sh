#!/bin/bash
# ncu captures (one gpurun call): launch list of a short decode + full-set capture of selected kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python tools/one_conv.py 1024 128 128 3 1 0 1 > gpurun_out/one_conv_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_conv128_res -f python tools/one_conv.py 1024 128 128 3 1 0 1 > gpurun_out/ncu_conv.log 2>&1
echo "ncu conv exit $?"
python tools/profile_decode.py 1 128 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_b1.csv python tools/profile_decode.py 1 128 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 40 -c 1 -o gpurun_out/prof_gn_apply -f python tools/profile_decode.py 1 128 > gpurun_out/ncu_gn.log 2>&1
echo "ncu gn exit $?"
ncu --set full --clock-control none --import-source on -k regex:hdr_phase_a -c 1 -o gpurun_out/prof_phase_a -f python tools/profile_decode.py 1 128 > gpurun_out/ncu_pa.log 2>&1
echo "ncu phase_a exit $?"
ls -la gpurun_out/*.ncu-rep
