# A/B of the two-CTA cluster latency kernel (TFHE_B200_CLUSTER=1) against K3L
mkdir -p gpurun_out
TFHE_B200_CLUSTER=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_circuits.py -m gpu -q -x 2>&1 | tail -3
: > gpurun_out/cluster_ab.jsonl
for v in ${VARIANTS:-0 1}; do echo "{\"cluster\": $v}" >> gpurun_out/cluster_ab.jsonl; TFHE_B200_CLUSTER=$v timeout 300 python tools/latency_probe.py >> gpurun_out/cluster_ab.jsonl 2>> gpurun_out/cluster_ab.err; done
cat gpurun_out/cluster_ab.jsonl; tail -5 gpurun_out/cluster_ab.err
