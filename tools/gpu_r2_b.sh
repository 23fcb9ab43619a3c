#!/bin/bash
# round 2, step B: fused attention — correctness first (bounded by timeout: a pipeline bug traps after seconds), then speed
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" ) > gpurun_out/r2b_attn_tests.log 2>&1
echo "attn tests rc=$?" | tee -a gpurun_out/r2b_attn_tests.log
tail -25 gpurun_out/r2b_attn_tests.log
for cfg in "1 2" "1 1" "0 -"; do
  set -- $cfg
  ( HDRVAE_ATTN_FUSED=$1 HDRVAE_ATTN_CG=$2 timeout 300 python tools/attn_bench.py ) >> gpurun_out/r2b_attn_bench.log 2>&1
  echo "bench fused=$1 cg=$2 rc=$?" >> gpurun_out/r2b_attn_bench.log
done
cat gpurun_out/r2b_attn_bench.log
( timeout 600 python -m pytest tests/test_gpu_parity_big.py tests/test_gpu_decode.py -m gpu -q -x ) > gpurun_out/r2b_decode_tests.log 2>&1
echo "decode tests rc=$?" | tee -a gpurun_out/r2b_decode_tests.log
tail -8 gpurun_out/r2b_decode_tests.log
