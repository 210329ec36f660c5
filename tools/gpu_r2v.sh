#!/bin/bash
# round 2, GPU call V: full-range parity (wide_check: chains vs layer-by-layer vs oracle; pytest) + iris (full mode) + c3 / c2 / c4 bench
mkdir -p gpurun_out
FDT_CHAIN=0 timeout 300 python tools/wide_check.py 2>&1 | tail -3
FDT_CHAIN=1 timeout 300 python tools/wide_check.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_mode.py -m gpu -q -x 2>&1 | tail -3
CFGS="${CFGS:-c3 c2 c4}" bash tools/gpu_r2r.sh | grep value
