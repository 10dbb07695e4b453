mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py 8 3 2 fused 2>&1 | grep "multi_gpu_check\|Error\|error" | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n4.log 2>&1; tail -1 gpurun_out/bench_n4.log | cut -c1-260
