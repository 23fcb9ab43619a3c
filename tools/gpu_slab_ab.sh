#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "128 0" "512 0" "512 1" "128 0" "512 0" "128 1"; do
  set -- $cfg
  HDRVAE_SLAB_MAXN=$1 HDRVAE_SLAB_RES=$2 python tools/graph_ab.py 2>&1 | tail -1 | sed "s/^/maxn=$1 res=$2: /"
done
HDRVAE_SLAB_MAXN=512 HDRVAE_SLAB_RES=0 timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -x --timeout 600 2>&1 | tail -2
HDRVAE_SLAB_MAXN=512 HDRVAE_SLAB_RES=0 python tools/profile_decode.py 4 128 gpurun_out/profile_c2_slab512.tsv | tail -1
