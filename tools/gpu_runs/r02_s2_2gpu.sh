#!/bin/bash
# Round 2, session 2: two GPUs -- the multi-GPU tests and the bench line with msplit + chain
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/s2_dist_tests_2gpu.log 2>&1; echo "dist tests rc=$?"; tail -4 gpurun_out/s2_dist_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/s2_bench_2gpu.json 2> gpurun_out/s2_bench_2gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/s2_bench_2gpu.err; head -c 700 gpurun_out/s2_bench_2gpu.json
