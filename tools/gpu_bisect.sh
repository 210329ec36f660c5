#!/bin/bash
mkdir -p gpurun_out
for d in 1 2 0; do
  echo "== FDT_STEM_TMTYPE=$d"; FDT_STEM_TMTYPE=$d timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
done
