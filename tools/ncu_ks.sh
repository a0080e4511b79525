#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline --unsplit"
$CMD > gpurun_out/plain_ks.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:keyswitch -s 1 -c 1 -f -o gpurun_out/prof_ks2 $CMD > gpurun_out/ncu_ks2.log 2>&1
ls -la gpurun_out/prof_ks2.ncu-rep
