#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2c_attn_bench.log
( HDRVAE_ATTN_FUSED=1 HDRVAE_ATTN_CG=2 timeout 300 python tools/attn_bench.py 1x16384 4x16384 4x16448 2x16384 1x32768 1x65536 4x4096 1x262144 ) >> gpurun_out/r2c_attn_bench.log 2>&1
( HDRVAE_ATTN_FUSED=1 HDRVAE_ATTN_CG=1 timeout 300 python tools/attn_bench.py 1x16384 4x16384 4x16448 ) >> gpurun_out/r2c_attn_bench.log 2>&1
cat gpurun_out/r2c_attn_bench.log
( timeout 600 python -m pytest tests/test_gpu_parity_big.py -m gpu -q -x -k "high" ) > gpurun_out/r2c_high.log 2>&1
echo "high rc=$?"; tail -15 gpurun_out/r2c_high.log | cut -c1-300
( timeout 600 python -m pytest tests/test_gpu_upscaler.py -m gpu -q -x ) > gpurun_out/r2c_up.log 2>&1
echo "upscaler rc=$?"; tail -5 gpurun_out/r2c_up.log | cut -c1-300
( timeout 600 python bench.py --steps 10 --warmup 3 --no-eager --no-cpu-baseline ) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?"; cut -c1-3500 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
