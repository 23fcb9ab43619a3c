#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 8 29551 tools/rows_bench.py 512 512 --check --steps 5 --mode aggressive > gpurun_out/rows_bench_n8.json 2> gpurun_out/rows_bench_n8.err; echo "rows n8 exit $?"; grep '^{' gpurun_out/rows_bench_n8.json
run 4 29553 tools/rows_bench.py 512 512 --steps 3 --mode aggressive > gpurun_out/rows_bench_n4.json 2> gpurun_out/rows_bench_n4.err; echo "rows n4 exit $?"; grep '^{' gpurun_out/rows_bench_n4.json
for n in 4 8; do
  run $n 2955$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n$n exit $?"
  python -c "
import json,sys
lines=[l for l in open('gpurun_out/bench_n$n.json')]
print('stdout lines', len(lines))
d=json.loads(lines[-1]); print('N=$n value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])"
done
