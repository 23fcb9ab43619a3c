#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/multi_test.log 2>&1; echo "multi test exit $?"; tail -15 gpurun_out/multi_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; tail -c 1500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
