#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2f_pytest.log | cut -c1-600
S="--steps 10 --warmup 3 --no-eager --no-cpu-baseline --no-aux"
for cfg in "1 2" "0 2" "1 1"; do
  set -- $cfg
  HDRVAE_FUSE_NIN=$1 HDRVAE_SILU_MUFU=$2 timeout 600 python bench.py $S > gpurun_out/r2f_bench_nin$1_silu$2.json 2> gpurun_out/r2f_bench.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2f_bench_nin$1_silu$2.json"))
print("nin=$1 silu=$2", round(d["value"], 2), "MP/s", round(d["ms_per_step"], 3), "ms", d["roofline"]["step_breakdown_ms"], d["clocks"]["sm_mhz"])
PY
done
python tools/profile_decode.py 4 128 gpurun_out/r2f_per_op_c2.tsv > gpurun_out/r2f_profile.log 2>&1; tail -1 gpurun_out/r2f_profile.log
