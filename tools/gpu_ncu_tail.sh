#!/bin/bash
# ncu --set full of one chunk (10 launches with the tail kernel) + SASS-level stall samples of k_tail_ws.  $1 = tag
TAG=${1:-x}
mkdir -p gpurun_out
python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 20 -c 10 -o /tmp/prof_$TAG -f python tools/prof_target.py > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>gpurun_out/ncu3.log
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:k_tail --launch-count 1 > gpurun_out/sass_k_tail_$TAG.csv 2>>gpurun_out/ncu3.log
ncu -i /tmp/prof_$TAG.ncu-rep --page details --csv -k regex:k_tail --launch-count 1 > gpurun_out/details_k_tail_$TAG.csv 2>>gpurun_out/ncu3.log
ls -la gpurun_out/ | head -30; tail -n 2 gpurun_out/ncu2.log; cat gpurun_out/prof_plain.log | tail -2
