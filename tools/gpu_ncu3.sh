#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full capture of the dominant kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_for_ncu.json 2> gpurun_out/bench_plain_for_ncu.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
python tools/profile_decode.py 4 128 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 20 -c 4 -o gpurun_out/prof_gemm_c2 -f python tools/profile_decode.py 4 128 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 25 -c 2 -o gpurun_out/prof_gn_c2 -f python tools/profile_decode.py 4 128 > gpurun_out/ncu_gn.log 2>&1
echo "ncu gn exit $?"
ncu --set full --clock-control none --import-source on -k regex:hdr_phase_a -c 1 -o gpurun_out/prof_phase_a_c2 -f python tools/profile_decode.py 4 128 > gpurun_out/ncu_pa.log 2>&1
echo "ncu phase a exit $?"
