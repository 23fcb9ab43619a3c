#!/bin/bash
# interleaved same-box A/B of step-level switches (the chip runs at its power cap: single runs drift by ~2 %)
mkdir -p gpurun_out
S="--steps 20 --warmup 5 --no-eager --no-cpu-baseline --no-aux"
: > gpurun_out/r2_ab.log
for rep in 1 2 3; do
  for cfg in "1 2 1" "0 2 1" "1 1 1" "1 2 0"; do
    set -- $cfg
    HDRVAE_FUSE_NIN=$1 HDRVAE_SILU_MUFU=$2 HDRVAE_H16=$3 timeout 600 python bench.py $S > gpurun_out/r2_ab_tmp.json 2> gpurun_out/r2_ab.err
    python - >> gpurun_out/r2_ab.log <<PY
import json
d = json.load(open("gpurun_out/r2_ab_tmp.json"))
print("rep $rep nin=$1 silu=$2 h16=$3", round(d["ms_per_step"], 3), "ms", round(d["value"], 2), "MP/s  clock", d["clocks"]["sm_mhz"], d["roofline"]["step_breakdown_ms"])
PY
  done
done
cat gpurun_out/r2_ab.log
