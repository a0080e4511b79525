timeout 900 python -m pytest tests -m gpu -q -x -k "mk or circuits" 2>&1 | tail -5
for f in 0 1; do for r in 1 0; do
FLAGS=$f TFHE_B200_MK_RING=$r timeout 600 python tools/mk_perf.py 2 2368
done; done
FLAGS=0 TFHE_B200_MK_RING=1 timeout 600 python tools/mk_perf.py 4 1184
FLAGS=0 TFHE_B200_MK_RING=0 timeout 600 python tools/mk_perf.py 4 1184
