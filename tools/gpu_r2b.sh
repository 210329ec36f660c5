#!/bin/bash
# round 2, GPU call B: whole GPU suite (no -x), tail kernel check
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -40 gpurun_out/pytest_gpu.log
FDT_TAIL=0 timeout 300 python tools/tail_check.py > gpurun_out/tail_check.log 2>&1; FDT_TAIL=1 timeout 300 python tools/tail_check.py >> gpurun_out/tail_check.log 2>&1; tail -30 gpurun_out/tail_check.log
FDT_TAIL=1 timeout 300 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2_tail.json 2> gpurun_out/bench_c2_tail.err; tail -c 400 gpurun_out/bench_c2_tail.err; head -c 300 gpurun_out/bench_c2_tail.json; echo
