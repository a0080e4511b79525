mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_device or cloud_key_over_all or dispatch_path or bootstrap_wo_ks" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
cat gpurun_out/bench_2gpu.json | cut -c1-3000; tail -3 gpurun_out/bench_2gpu.err
