#!/bin/bash
# per-instruction executed counts of one launch inside a bench.py run: gpu_ncu_src3.sh <config> <kernel regex> <skip> <tag>
mkdir -p gpurun_out
c=$1; KN=$2; skip=$3; tag=$4
ncu --clock-control none --section SourceCounters --section WarpStateStats --section SpeedOfLight --import-source on -k regex:$KN -s $skip -c 1 -o /tmp/src_$tag -f \
    python bench.py --config $c --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_src_$tag.log 2>&1
ncu -i /tmp/src_$tag.ncu-rep --page source --csv --print-source sass > gpurun_out/src_$tag.csv 2>> gpurun_out/ncu_src_$tag.log
ncu -i /tmp/src_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>> gpurun_out/ncu_src_$tag.log
tail -2 gpurun_out/ncu_src_$tag.log
