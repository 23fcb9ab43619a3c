#!/bin/bash
# One gpurun call: parity tests (each file under its own timeout so a hang cannot eat the box), smoke, profile, bench.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rm -f gpurun_out/summary.txt
for f in test_gpu_kernels test_gpu_epilogue test_gpu_decode test_gpu_upscaler; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  tail -4 gpurun_out/$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
timeout 300 python tools/profile_decode.py 4 128 gpurun_out/profile_c2.tsv > gpurun_out/profile_c2.log 2>&1; echo "profile exit $?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/profile_c2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt; tail -2 gpurun_out/bench.json
