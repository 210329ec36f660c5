#!/bin/bash
# per-instruction executed counts / stall samples of one k_block_ts launch (variant $1, skip $2 launches of k_block_ts)
mkdir -p gpurun_out
v=$1; skip=${2:-12}; KN=${3:-k_block_ts}
export FDT_CUDA_LIB=$PWD/variants/lib_$v.so
python tools/prof_target.py 1024 2>&1 | tail -1
ncu --clock-control none --section SourceCounters --section WarpStateStats --import-source on -k regex:$KN -s $skip -c 1 -o /tmp/src_${v}_$KN -f python tools/prof_target.py 1024 > gpurun_out/ncu_src_${v}_$KN.log 2>&1
ncu -i /tmp/src_${v}_$KN.ncu-rep --page source --csv --print-source sass > gpurun_out/src_${v}_$KN.csv 2>> gpurun_out/ncu_src_${v}_$KN.log
