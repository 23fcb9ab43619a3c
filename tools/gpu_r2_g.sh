#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log | cut -c1-600
python tools/profile_decode.py 4 128 gpurun_out/r2g_per_op_c2.tsv > gpurun_out/r2g_profile.log 2>&1; tail -1 gpurun_out/r2g_profile.log
grep "128->128\|256->256 @4x512\|512->512 @4x256x256" gpurun_out/r2g_per_op_c2.tsv
S="--steps 20 --warmup 5 --no-eager --no-cpu-baseline --no-aux"
for rep in 1 2; do
  timeout 600 python bench.py $S > gpurun_out/r2g_bench_$rep.json 2> gpurun_out/r2g_bench.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2g_bench_$rep.json"))
print("rep $rep", round(d["ms_per_step"], 3), "ms", round(d["value"], 2), "MP/s  clock", d["clocks"]["sm_mhz"], d["roofline"]["step_breakdown_ms"])
PY
done
