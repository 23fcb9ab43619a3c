#!/bin/bash
# last check of the round on the in-tree build: whole GPU suite, smoke(), the default bench line
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/last_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/last_pytest.log | cut -c1-300
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/last_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/last_smoke.log | cut -c1-300
( timeout 900 python bench.py ) > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
echo "bench rc=$?"; cut -c1-260 gpurun_out/last_bench.json
