mkdir -p gpurun_out
for g in ${PARITY:-354 364}; do TFHE_B200_G=$g timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2; done
: > gpurun_out/exp_perf.jsonl
for g in ${VARIANTS:-304 354 364}; do TFHE_B200_G=$g timeout 300 python tools/quick_perf.py ${B:-65536} >> gpurun_out/exp_perf.jsonl 2>> gpurun_out/exp_perf.err; done
cat gpurun_out/exp_perf.jsonl
for g in ${PROBES:-1334}; do TFHE_B200_G=$g timeout 200 python tools/quick_perf.py 2368 2>&1 | grep -E "probe cta 0" | head -8; done
