#!/bin/bash
# One ncu pass: per-launch duration + DRAM bytes of every kernel of C2 decodes (graphs off so each kernel is a launch).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export HDRVAE_NO_GRAPH=1
python tools/profile_decode.py 4 128 > gpurun_out/r2_traffic_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_traffic_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv \
    --log-file gpurun_out/r2_traffic_c2.csv python tools/profile_decode.py 4 128 > gpurun_out/r2_traffic_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r2_traffic_ncu.log; wc -l gpurun_out/r2_traffic_c2.csv
