#!/bin/bash
# HDRVAE_X16 (residual stream as a scaled 16-bit tensor): accuracy sweep with and without, decode tests with it, interleaved A/B
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decode.py -m gpu -q -x ) > gpurun_out/x16_pytest_default.log 2>&1
echo "pytest (default) rc=$?"; tail -2 gpurun_out/x16_pytest_default.log | cut -c1-300
HDRVAE_X16=0 timeout 600 python tools/parity_sweep.py > gpurun_out/x16_sweep_off.txt 2>&1
HDRVAE_X16=1 timeout 600 python tools/parity_sweep.py > gpurun_out/x16_sweep_on.txt 2>&1
echo "--- off"; cat gpurun_out/x16_sweep_off.txt | tail -13; echo "--- on"; tail -13 gpurun_out/x16_sweep_on.txt
( HDRVAE_X16=1 timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_parity_big.py -m gpu -q ) > gpurun_out/x16_pytest_on.log 2>&1
echo "pytest (x16) rc=$?"; tail -8 gpurun_out/x16_pytest_on.log | cut -c1-300
S="--steps 20 --warmup 5 --no-eager --no-cpu-baseline --no-aux"
: > gpurun_out/x16_ab.log
for rep in 1 2 3; do
  for x in 0 1; do
    HDRVAE_X16=$x timeout 600 python bench.py $S > gpurun_out/x16_tmp.json 2> gpurun_out/x16_tmp.err
    python - >> gpurun_out/x16_ab.log <<PY
import json
d = json.load(open("gpurun_out/x16_tmp.json"))
print("rep $rep x16=$x", round(d["ms_per_step"], 3), "ms", round(d["value"], 2), "MP/s  clock", d["clocks"]["sm_mhz"], d["roofline"]["step_breakdown_ms"])
PY
  done
done
cat gpurun_out/x16_ab.log
HDRVAE_X16=1 python tools/profile_decode.py 4 128 gpurun_out/x16_per_op.tsv > gpurun_out/x16_profile.log 2>&1; tail -1 gpurun_out/x16_profile.log
