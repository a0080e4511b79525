#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), unsplit A/B, MK / latency / circuit probes, ncu launch list of
# the bench command and one full capture of each dominant kernel (ncu only after the same command exited 0 without it).
# Everything lands in gpurun_out/ (keep it under 64 MiB).
set -u
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
fi
timeout 900 python bench.py > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
FLAGS=1 timeout 300 python tools/quick_perf.py 65536 > gpurun_out/quick_unsplit.json 2> gpurun_out/quick_unsplit.err
timeout 300 python tools/quick_perf.py 65536 > gpurun_out/quick_split.json 2> gpurun_out/quick_split.err
if [ "${EXTRA:-1}" = "1" ]; then
  timeout 300 python tools/circuit_latency.py > gpurun_out/circuits_split.json 2> gpurun_out/circuits.err
  timeout 300 python tools/latency_probe.py > gpurun_out/latency.json 2> gpurun_out/latency.err
  timeout 300 python tools/mk_perf.py 2 2368 > gpurun_out/mk_perf_p2.json 2> gpurun_out/mk_perf.err
  SAMPLE=4 timeout 300 python tools/mk_perf.py 4 1184 > gpurun_out/mk_perf_p4.json 2>> gpurun_out/mk_perf.err
  timeout 300 python tools/latency128.py > gpurun_out/latency128.json 2> gpurun_out/latency128.err
  timeout 900 python tools/sweep.py > gpurun_out/sweep.json 2> gpurun_out/sweep.err
  timeout 300 python tools/mask_size_perf.py > gpurun_out/mask_size.json 2> gpurun_out/mask_size.err
fi
NCU="ncu --set full --import-source on --clock-control none -f"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"   # the default workload (2^17 gates per launch)
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && $NCU -k regex:blind_rotate_kernel -s 1 -c 1 -o gpurun_out/prof_blind_rotate $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log 2>/dev/null; cat gpurun_out/smoke.log 2>/dev/null
cat gpurun_out/bench_split.json gpurun_out/bench_reference.json | cut -c1-1500
cat gpurun_out/quick_split.json gpurun_out/quick_unsplit.json gpurun_out/circuits_split.json gpurun_out/latency.json gpurun_out/mk_perf_p2.json gpurun_out/sweep.json 2>/dev/null
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
