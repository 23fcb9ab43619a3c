#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of the bench command, full captures of the fused attention kernel and
# of the GroupNorm apply kernel.  Each ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager --no-aux"
$B > gpurun_out/r2_ncu_plain_bench.json 2> gpurun_out/r2_ncu_plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_ncu_bench.log 2>&1
echo "ncu launch list exit $?"
A="python tools/attn_bench.py 1x32768"
$A > gpurun_out/r2_ncu_plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 1 -c 1 -o gpurun_out/r2_prof_attn_fused -f $A > gpurun_out/r2_ncu_attn.log 2>&1
echo "ncu attn exit $?"
P="python tools/profile_decode.py 4 128"
$P > gpurun_out/r2_ncu_plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 84 -c 3 -o gpurun_out/r2_prof_gn_apply -f $P > gpurun_out/r2_ncu_gn.log 2>&1
echo "ncu gn exit $?"
ls -la gpurun_out/*.ncu-rep
