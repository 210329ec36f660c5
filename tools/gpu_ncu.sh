#!/bin/bash
# ncu evidence: launch list of one chunk + full-set capture of its kernels, exported as CSV on the box
# (the .ncu-rep itself is too large to bring back).  $1 = tag
TAG=${1:-r1}
mkdir -p gpurun_out
python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 46 -c 23 --csv --log-file gpurun_out/launches_$TAG.csv python tools/prof_target.py > gpurun_out/ncu1.log 2>&1
python tools/prof_target.py > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_dwpw|k_gemm_conv|k_letterbox|k_decode' -s 46 -c 23 -o /tmp/prof_$TAG -f python tools/prof_target.py > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>gpurun_out/ncu3.log
ncu -i /tmp/prof_$TAG.ncu-rep --page details --csv > gpurun_out/details_$TAG.csv 2>>gpurun_out/ncu3.log
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:k_gemm_conv > gpurun_out/source_gemm_$TAG.csv 2>>gpurun_out/ncu3.log
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:k_dwpw --launch-skip 1 --launch-count 1 > gpurun_out/source_dwpw2_$TAG.csv 2>>gpurun_out/ncu3.log
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:k_dwpw --launch-skip 12 --launch-count 1 > gpurun_out/source_dwpw13_$TAG.csv 2>>gpurun_out/ncu3.log
tail -n 3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log; ls -la gpurun_out/ /tmp/*.ncu-rep
