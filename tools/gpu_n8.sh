#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/rows_bench.py 512 512 --check --steps 5 --mode aggressive > gpurun_out/rows_bench_n8.json 2> gpurun_out/rows_bench_n8.err; echo "rows n8 exit $?"; cat gpurun_out/rows_bench_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json')); print('N=8 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])"
tail -3 gpurun_out/bench_n8.err
