#!/bin/bash
# round 2, GPU call V: full-range parity (wide_check: chains vs layer-by-layer vs oracle; pytest full-range cases) + c3 bench
mkdir -p gpurun_out
FDT_CHAIN=0 timeout 300 python tools/wide_check.py 2>&1 | tail -3
FDT_CHAIN=1 timeout 300 python tools/wide_check.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full or kernel_variants" 2>&1 | tail -3
CFGS="c3" bash tools/gpu_r2r.sh | grep value
