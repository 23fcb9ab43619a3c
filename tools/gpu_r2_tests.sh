#!/bin/bash
# round 2, step A: full GPU test suite + default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print(os.cpu_count())" >> gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -30 gpurun_out/r2t_pytest.log
echo skip bench
echo "bench rc=$?"


