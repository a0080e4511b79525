#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/latency_probe.py > gpurun_out/latency.json 2> gpurun_out/latency.err; cat gpurun_out/latency.json; tail -2 gpurun_out/latency.err
CPU=0 timeout 300 python tools/circuit_latency.py | tee gpurun_out/circuits_split.json; FLAGS=1 CPU=0 timeout 300 python tools/circuit_latency.py | tee gpurun_out/circuits_unsplit.json
