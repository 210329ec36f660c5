#!/bin/bash
# round 2, GPU call T: stem with the im2col operand in TMEM: parity (every materialised tensor incl. the stem output, mesh / iris stems) + benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_full_mode.py -m gpu -q -x 2>&1 | tail -3
CFGS="c2 c3 c4" bash tools/gpu_r2r.sh | grep value
