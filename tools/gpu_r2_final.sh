#!/bin/bash
# Round-2 closing evidence on one GPU: full GPU suite, bench, per-op table, issuer wait table, then the ncu passes
# (launch list of the bench command, DRAM traffic per kernel, full captures of two conv launches) — each ncu pass after the
# same command has exited 0 without ncu.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -12 gpurun_out/r2f_pytest.log | cut -c1-300
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; cut -c1-220 gpurun_out/r2f_bench.json; tail -2 gpurun_out/r2f_bench.err
python tools/profile_decode.py 4 128 gpurun_out/r2f_per_op_c2.tsv > gpurun_out/r2f_profile.log 2>&1; tail -1 gpurun_out/r2f_profile.log
HDRVAE_NO_GRAPH=1 HDRVAE_GEMM_DBG=32 timeout 300 python tools/profile_decode.py 4 128 > gpurun_out/r2f_dbg32_all.log 2>&1
grep "gemm_tc<" gpurun_out/r2f_dbg32_all.log | tail -124 > gpurun_out/r2f_dbg32.log; wc -l gpurun_out/r2f_dbg32.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager --no-aux"
$B > gpurun_out/r2f_ncu_plain_bench.json 2> gpurun_out/r2f_ncu_plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches_bench.csv $B > gpurun_out/r2f_ncu_bench.log 2>&1
echo "ncu launch list exit $?"
export HDRVAE_NO_GRAPH=1
python tools/profile_decode.py 4 128 > gpurun_out/r2f_traffic_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/r2f_traffic_c2.csv python tools/profile_decode.py 4 128 > gpurun_out/r2f_traffic_ncu.log 2>&1
echo "ncu traffic exit $?"
unset HDRVAE_NO_GRAPH
A="python tools/one_conv.py 1024 128 128 3 1 0 4"
$A > gpurun_out/r2f_ncu_plain_res.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r2f_prof_conv128_res -f $A > gpurun_out/r2f_ncu_res.log 2>&1
echo "ncu res exit $?"
C="python tools/one_conv.py 1024 128 128 3 0 0 4"
$C > gpurun_out/r2f_ncu_plain_nores.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r2f_prof_conv128_nores -f $C > gpurun_out/r2f_ncu_nores.log 2>&1
echo "ncu nores exit $?"
ls -la gpurun_out/r2f_prof_conv128*.ncu-rep
