#!/bin/bash
# Round 2, 8 GPUs of one box: D2H ceiling, the m-split tests, and the bench line with the msplit and
# chain keys.   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_runs/r02_8gpu.sh'
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 tools/probe_d2h.py 2>/dev/null | tee gpurun_out/r02_d2h_8gpu.json
$TR --nproc-per-node 4 --master-port 29522 tools/probe_d2h.py 2>/dev/null | tee gpurun_out/r02_d2h_4gpu.json
$TR --nproc-per-node 2 --master-port 29523 tools/probe_d2h.py 2>/dev/null | tee gpurun_out/r02_d2h_2gpu.json
python tools/probe_d2h.py | tee gpurun_out/r02_d2h_1gpu.json
$TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/r02_bench_8gpu.err | tee gpurun_out/r02_bench_8gpu.json | cut -c1-300
tail -3 gpurun_out/r02_bench_8gpu.err
timeout 300 $TR --nproc-per-node 8 --master-port 29525 tests/dist/msplit_check.py --p2p 1024 4096 2>&1 | grep -v "^W\|^\*" | tee gpurun_out/r02_msplit_8gpu.log
python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_dist_tests_8gpu.log
