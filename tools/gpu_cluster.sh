# A/B of the two-CTA cluster latency kernel against K3L.  TFHE_B200_CLUSTER: 0 off, 1 keys through the TMA slots,
# 3 keys straight into registers; 2 / 4 print the clock64 phase probe of 1 / 3.
mkdir -p gpurun_out
for v in ${PARITY:-3}; do TFHE_B200_CLUSTER=$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_circuits.py -m gpu -q -x 2>&1 | tail -3; done
: > gpurun_out/cluster_probe.txt
for v in ${PROBES:-2 4}; do echo "== TFHE_B200_CLUSTER=$v" >> gpurun_out/cluster_probe.txt; TFHE_B200_CLUSTER=$v timeout 200 python tools/latency_probe.py 2>&1 | grep "cluster probe" | head -8 >> gpurun_out/cluster_probe.txt; done
cat gpurun_out/cluster_probe.txt
: > gpurun_out/cluster_ab.jsonl
for v in ${VARIANTS:-0 1 3}; do echo "{\"cluster\": $v}" >> gpurun_out/cluster_ab.jsonl; TFHE_B200_CLUSTER=$v timeout 300 python tools/latency_probe.py >> gpurun_out/cluster_ab.jsonl 2>> gpurun_out/cluster_ab.err; done
cat gpurun_out/cluster_ab.jsonl; tail -5 gpurun_out/cluster_ab.err
