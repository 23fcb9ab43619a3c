#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python tools/up_bench.py 1024 1024 23 gpurun_out/profile_up_c5.tsv > gpurun_out/up_bench.log 2>&1; echo "up bench exit $?"; tail -5 gpurun_out/up_bench.log
