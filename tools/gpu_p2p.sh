#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}; H=${2:-256}
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/multi_test.log 2>&1; echo "multi test exit $?"; tail -4 gpurun_out/multi_test.log
for tr in nccl p2p; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 tools/rows_bench.py $H $H --steps 10 --transport $tr 2> gpurun_out/rows_$tr.err | grep '^{' | tee gpurun_out/rows_${tr}_n$N.json
done
