#!/bin/bash
# usage: gpu_rows_n.sh N h w [extra args]   (run under gpurun --gpus N)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=$1; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/rows_bench.py "$@" > gpurun_out/rows_bench_n$N.json 2> gpurun_out/rows_bench_n$N.err; echo "rows bench exit $?"; cat gpurun_out/rows_bench_n$N.json; tail -5 gpurun_out/rows_bench_n$N.err
