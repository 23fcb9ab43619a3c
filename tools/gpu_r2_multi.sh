#!/bin/bash
# N-GPU bench (headline + multi_gpu_selftest + aux_c4_rows), then the real-NCCL pytest file.  usage: gpu_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench N=$N rc=$?"; cut -c1-300 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n$N.json"))
    for k in ("value", "ms_per_step", "e2e", "multi_gpu_selftest", "aux_c4_rows"):
        print(k, d.get(k))
except Exception as e:
    print("parse failed", e)
PY
if [ "$N" = "2" ]; then
  ( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r2_multi_tests.log 2>&1
  echo "multi tests rc=$?"; tail -5 gpurun_out/r2_multi_tests.log | cut -c1-300
fi
