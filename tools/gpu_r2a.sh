#!/bin/bash
# round 2, GPU call A: TS-MMA probe, the whole GPU test-suite, smoke, bench lines for c2 / c3 / c4
mkdir -p gpurun_out
(cd tools/probe && nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/ts_probe ts_probe.cu && /tmp/ts_probe; nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tsh_probe tsh_probe.cu && /tmp/tsh_probe) > gpurun_out/ts_probe.log 2>&1
cat gpurun_out/ts_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
for c in c2 c3 c4; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -c 600 gpurun_out/bench_$c.err; head -c 700 gpurun_out/bench_$c.json; echo
done
FDT_TAIL=0 timeout 300 python tools/tail_check.py > gpurun_out/tail_check.log 2>&1; FDT_TAIL=1 timeout 300 python tools/tail_check.py >> gpurun_out/tail_check.log 2>&1; tail -30 gpurun_out/tail_check.log
