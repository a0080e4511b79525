#!/bin/bash
# ncu --set full captures of the blind-rotation kernel, split and unsplit builds (4096 gates each).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline"
$CMD > gpurun_out/plain_s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -f -o gpurun_out/prof_br_split $CMD > gpurun_out/ncu_s.log 2>&1
$CMD --unsplit > gpurun_out/plain_u.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -f -o gpurun_out/prof_br_unsplit $CMD --unsplit > gpurun_out/ncu_u.log 2>&1
ls -la gpurun_out/*.ncu-rep
