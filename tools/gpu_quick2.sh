#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for f in test_gpu_epilogue test_gpu_decode; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; tail -3 gpurun_out/$f.log
done
HDRVAE_TC_CONVOUT=0 timeout 300 python tools/profile_decode.py 4 128 gpurun_out/profile_c2_cc.tsv > gpurun_out/profile_c2_cc.log 2>&1; tail -1 gpurun_out/profile_c2_cc.log; grep "epilogue" gpurun_out/profile_c2_cc.tsv | cut -c1-100
timeout 300 python tools/profile_decode.py 4 128 gpurun_out/profile_c2.tsv > gpurun_out/profile_c2.log 2>&1; tail -1 gpurun_out/profile_c2.log; grep "epilogue\|128->8" gpurun_out/profile_c2.tsv | cut -c1-100
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
