# A/B of the latency kernels: TFHE_B200_LOWLAT = 1 (K3L), 2 (K3L2, keys through the TMA ring), 3 (K3L2, keys through TMEM)
mkdir -p gpurun_out
for v in ${PARITY:-2 3}; do TFHE_B200_LOWLAT=$v timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_circuits.py -m gpu -q -x 2>&1 | tail -3; done
: > gpurun_out/lowlat_ab.jsonl
for v in ${VARIANTS:-1 2 3}; do echo "{\"lowlat\": $v}" >> gpurun_out/lowlat_ab.jsonl; TFHE_B200_LOWLAT=$v timeout 300 python tools/latency_probe.py >> gpurun_out/lowlat_ab.jsonl 2>> gpurun_out/lowlat_ab.err; done
cat gpurun_out/lowlat_ab.jsonl
