mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python tools/latency_probe.py > gpurun_out/latency.json 2> gpurun_out/latency.err
timeout 300 python tools/circuit_latency.py > gpurun_out/circuits_split.json 2> gpurun_out/circuits.err
TFHE_B200_CLUSTER=2 timeout 200 python tools/latency_probe.py 2>&1 | grep "cluster probe" | head -8 > gpurun_out/cluster_probe.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/latency.json gpurun_out/circuits_split.json gpurun_out/cluster_probe.txt; cut -c1-600 gpurun_out/bench_split.json
