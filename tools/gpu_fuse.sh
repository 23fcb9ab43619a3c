#!/bin/bash
# fused GroupNorm + SiLU operand transform (HDRVAE_FUSE_GN=1): parity tests, then same-box A/B of the C2 step
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
HDRVAE_FUSE_GN=1 timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -x --timeout 300 > gpurun_out/fuse_test.log 2>&1; echo "fused decode tests exit $?"; tail -6 gpurun_out/fuse_test.log | cut -c1-300
for f in 0 1 0 1; do HDRVAE_FUSE_GN=$f timeout 200 python tools/graph_ab.py 2>&1 | tail -1 | sed "s/^/fuse=$f: /"; done
HDRVAE_FUSE_GN=1 timeout 200 python tools/profile_decode.py 4 128 gpurun_out/profile_c2_fused.tsv 2>&1 | tail -1
