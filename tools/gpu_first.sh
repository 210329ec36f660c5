#!/bin/bash
# First GPU contact: diagnostics (per-tensor parity table), the GPU test-suite, a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 900 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1; echo "diag exit $?" >> gpurun_out/diag.log
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/diag.log; tail -15 gpurun_out/pytest_gpu.log; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
