#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python tools/profile_decode.py 1 128 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hdr_phase_a -c 1 -o gpurun_out/prof_phase_a_v2 -f python tools/profile_decode.py 1 128 > gpurun_out/ncu_pa.log 2>&1
echo "ncu phase a exit $?"
