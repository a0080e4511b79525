mkdir -p gpurun_out
: > gpurun_out/lowlat_probe.txt
for v in ${PROBES:-12 13}; do echo "== TFHE_B200_LOWLAT=$v" >> gpurun_out/lowlat_probe.txt; TFHE_B200_LOWLAT=$v timeout 200 python tools/quick_perf.py 256 2>&1 | grep -E "lowlat2 probe" | head -8 >> gpurun_out/lowlat_probe.txt; done
cat gpurun_out/lowlat_probe.txt
