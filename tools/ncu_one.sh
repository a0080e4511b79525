#!/bin/bash
# usage: tools/ncu_one.sh <tag> [bench args...]   (env such as TFHE_B200_G is inherited)
set -u
TAG=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
