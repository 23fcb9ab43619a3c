#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decode.py -m gpu -q -x ) > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log | cut -c1-600
python tools/profile_decode.py 4 128 gpurun_out/r2h_per_op_c2.tsv > gpurun_out/r2h_profile.log 2>&1; tail -1 gpurun_out/r2h_profile.log
grep groupnorm gpurun_out/r2h_per_op_c2.tsv | awk -F'\t' '{split($2,a," "); s+=a[1]} END {print "GN total (burst clocks):", s}'
grep "groupnorm" gpurun_out/r2h_per_op_c2.tsv | tail -8
grep "128->128" gpurun_out/r2h_per_op_c2.tsv
S="--steps 20 --warmup 5 --no-eager --no-cpu-baseline --no-aux"
for rep in 1 2 3; do
  timeout 600 python bench.py $S > gpurun_out/r2h_bench_$rep.json 2> gpurun_out/r2h_bench.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2h_bench_$rep.json"))
print("rep $rep", round(d["ms_per_step"], 3), "ms", round(d["value"], 2), "MP/s  clock", d["clocks"]["sm_mhz"], d["roofline"]["step_breakdown_ms"])
PY
done
