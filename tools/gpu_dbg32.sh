#!/bin/bash
# where do the conv kernel's issuer / epilogue / transform warps wait?  (HDRVAE_GEMM_DBG=32: CTA 0 prints its wait cycles per launch)
mkdir -p gpurun_out
HDRVAE_NO_GRAPH=1 HDRVAE_GEMM_DBG=32 timeout 300 python tools/profile_decode.py 4 128 > gpurun_out/dbg32.log 2>&1
grep -c "gemm_tc<" gpurun_out/dbg32.log
grep "gemm_tc<" gpurun_out/dbg32.log | tail -190 > gpurun_out/dbg32_last.log
grep "transform" gpurun_out/dbg32_last.log | tail -12
