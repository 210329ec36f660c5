#!/bin/bash
# round 2, GPU call F: ncu (one chunk, all kernels + SASS stalls of k_tail_ws), then the whole GPU suite with the new defaults
bash tools/gpu_ncu_tail.sh r02c > gpurun_out/ncu_tail.log 2>&1; tail -3 gpurun_out/ncu_tail.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -30 gpurun_out/pytest_gpu.log
