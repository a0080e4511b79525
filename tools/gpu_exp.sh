#!/bin/bash
# Development GPU session: A/B of blind-rotation kernel variants + parity of the variant under test.
# Usage: gpurun -- 'bash tools/gpu_exp.sh "304 0" 65536'     (variants = values of TFHE_B200_G; 0 = library default)
set -u
mkdir -p gpurun_out
VARIANTS=${1:-"0 304"}
B=${2:-65536}
: > gpurun_out/exp_perf.jsonl
for g in $VARIANTS; do
  TFHE_B200_G=$g timeout 300 python tools/quick_perf.py $B >> gpurun_out/exp_perf.jsonl 2>> gpurun_out/exp_perf.err
done
cat gpurun_out/exp_perf.jsonl
for g in $VARIANTS; do
  if [ "$g" != "0" ]; then
    TFHE_B200_G=$g timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_G$g.log 2>&1
    echo "G=$g parity exit $?"; tail -3 gpurun_out/pytest_G$g.log
  fi
done
if [ "${RUN_SUITE:-1}" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -rA --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "suite exit $?"
  grep -E "passed|failed|error|parties:|^[0-9.]+s " gpurun_out/pytest_gpu.log | tail -25
fi
if [ -n "${NCU_G:-}" ]; then
  TFHE_B200_G=$NCU_G ncu --set full --import-source on --clock-control none -f -k regex:blind_rotate_kernel -s 1 -c 1 \
     -o gpurun_out/prof_br_G$NCU_G python tools/quick_perf.py ${NCU_B:-16384} > gpurun_out/ncu_G$NCU_G.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
