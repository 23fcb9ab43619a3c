#!/bin/bash
# same-box A/B of two builds of libhdrvae.so: tools/ab/libhdrvae_base.so (previous) against the in-tree one (new).
# Runs on the GPU box's scratch copy of the repo, so swapping the in-tree file there is harmless.
mkdir -p gpurun_out
L=vae_decode_hdr_b200/libhdrvae.so
cp $L /tmp/libhdrvae_new.so
( timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/ab_pytest.log 2>&1
echo "pytest(new) rc=$?"; tail -3 gpurun_out/ab_pytest.log | cut -c1-400
S="--steps 20 --warmup 5 --no-eager --no-cpu-baseline --no-aux"
: > gpurun_out/ab_so.log
for rep in 1 2 3; do
  for which in base new; do
    if [ $which = base ]; then cp tools/ab/libhdrvae_base.so $L; else cp /tmp/libhdrvae_new.so $L; fi
    timeout 600 python bench.py $S > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
    python - >> gpurun_out/ab_so.log <<PY
import json
d = json.load(open("gpurun_out/ab_tmp.json"))
print("rep $rep $which", round(d["ms_per_step"], 3), "ms", round(d["value"], 2), "MP/s  clock", d["clocks"]["sm_mhz"], d["roofline"]["step_breakdown_ms"])
PY
  done
done
cat gpurun_out/ab_so.log
cp tools/ab/libhdrvae_base.so $L; python tools/profile_decode.py 4 128 gpurun_out/ab_per_op_base.tsv > gpurun_out/ab_profile_base.log 2>&1; tail -1 gpurun_out/ab_profile_base.log
cp /tmp/libhdrvae_new.so $L; python tools/profile_decode.py 4 128 gpurun_out/ab_per_op_new.tsv > gpurun_out/ab_profile_new.log 2>&1; tail -1 gpurun_out/ab_profile_new.log
paste gpurun_out/ab_per_op_base.tsv gpurun_out/ab_per_op_new.tsv | awk -F'\t' '{printf "%-50s %s -> %s\n", $1, $2, $6}' | grep -v groupnorm
HDRVAE_NO_GRAPH=1 HDRVAE_GEMM_DBG=32 timeout 300 python tools/profile_decode.py 4 128 > gpurun_out/dbg32.log 2>&1
grep "gemm_tc<" gpurun_out/dbg32.log | tail -150 > gpurun_out/dbg32_last.log
# the 4x HDR upscaler (narrow 32- / 64-column convs), previous build then new build
cp tools/ab/libhdrvae_base.so $L; timeout 600 python tools/up_bench.py 1024 1024 23 > gpurun_out/ab_up_base.log 2>&1; tail -1 gpurun_out/ab_up_base.log
cp /tmp/libhdrvae_new.so $L; timeout 600 python tools/up_bench.py 1024 1024 23 > gpurun_out/ab_up_new.log 2>&1; tail -1 gpurun_out/ab_up_new.log
