# Persisting-L2 access-policy window over the key: DRAM bytes and time of one K3 launch (65 536 gates)
mkdir -p gpurun_out
: > gpurun_out/l2persist.txt
for v in 0 60 100; do
  echo "== TFHE_B200_L2PERSIST=$v" >> gpurun_out/l2persist.txt
  TFHE_B200_VERBOSE=1 TFHE_B200_L2PERSIST=$v timeout 300 python tools/quick_perf.py 65536 >> gpurun_out/l2persist.txt 2>&1
  TFHE_B200_L2PERSIST=$v timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:blind_rotate_kernel -s 1 -c 1 python tools/quick_perf.py 65536 2>&1 | grep -E "dram__|gpu__time|lts__|blind_rotate_kernel" >> gpurun_out/l2persist.txt
done
cat gpurun_out/l2persist.txt
