mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mk.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
: > gpurun_out/mk_ab.jsonl
for pw in 0 1; do
  TFHE_B200_MK_PW=$pw SAMPLE=4 timeout 300 python tools/mk_perf.py 2 2368 >> gpurun_out/mk_ab.jsonl 2>> gpurun_out/mk_ab.err
  TFHE_B200_MK_PW=$pw SAMPLE=2 timeout 300 python tools/mk_perf.py 4 1184 >> gpurun_out/mk_ab.jsonl 2>> gpurun_out/mk_ab.err
done
cat gpurun_out/mk_ab.jsonl
FLAGS=1 timeout 300 python tools/quick_perf.py 65536
