#!/bin/bash
# Round 2, session 2, first call (one GPU): gpurun --timeout 1700 -- 'bash tools/gpu_runs/r02_s2_first.sh'
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s2_gputests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/s2_gputests.log
timeout 500 python bench.py > gpurun_out/s2_bench_1gpu.json 2> gpurun_out/s2_bench_1gpu.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/s2_bench_1gpu.json
timeout 1000 bash tools/gpu_runs/prof_r02.sh > gpurun_out/s2_prof.log 2>&1; echo "prof rc=$?"
