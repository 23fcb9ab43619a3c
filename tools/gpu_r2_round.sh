#!/bin/bash
# full GPU suite + bench (with the same-GPU eager comparator and cpu baseline) + per-op profile
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -16 gpurun_out/r2d_pytest.log | cut -c1-400
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
( HDRVAE_H16=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-eager --no-cpu-baseline --no-aux ) > gpurun_out/r2d_bench_h32.json 2> gpurun_out/r2d_bench_h32.err
echo "bench h32 rc=$?"; cut -c1-200 gpurun_out/r2d_bench_h32.json
python tools/profile_decode.py 4 128 gpurun_out/r2d_per_op_c2.tsv > gpurun_out/r2d_profile.log 2>&1
tail -3 gpurun_out/r2d_profile.log
