# L2 eviction-policy hint on the key stream: DRAM bytes and time of one K3 launch (65 536 gates), hint off / on
mkdir -p gpurun_out
: > gpurun_out/l2hint.txt
for v in 0 1; do
  echo "== TFHE_B200_L2HINT=$v" >> gpurun_out/l2hint.txt
  TFHE_B200_L2HINT=$v timeout 300 python tools/quick_perf.py 65536 >> gpurun_out/l2hint.txt 2>&1
  TFHE_B200_L2HINT=$v timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:blind_rotate_kernel -s 1 -c 1 python tools/quick_perf.py 65536 2>&1 | grep -E "dram__|gpu__time|lts__|blind_rotate_kernel" >> gpurun_out/l2hint.txt
done
cat gpurun_out/l2hint.txt
