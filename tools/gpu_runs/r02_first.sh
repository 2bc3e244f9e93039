#!/bin/bash
# Round 2, first 1-GPU call: everything written at the end of round 1 without a GPU.
#   gpurun --timeout 1500 -- 'bash tools/gpu_runs/r02_first.sh'
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_tests.log
python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/r02_bench_reference.json
python bench.py | tee gpurun_out/r02_bench_1gpu.json
python tools/probe_solver.py 60 8191 3 | tee gpurun_out/r02_solver_ncorr3.json
python tools/probe_solver.py 60 8191 59 | tee gpurun_out/r02_solver_full.json
# K12 with 64-byte (default) and 32-byte L2 fetches
python tools/probe_sampling.py 2>&1 | tail -4 | tee gpurun_out/r02_sampling_default.log
GLB_L2_FETCH_BYTES=32 python tools/probe_sampling.py 2>&1 | tail -4 | tee gpurun_out/r02_sampling_fetch32.log
# device-walked batch cuts (opt-in): the points suite with the knob on
GLB_POINTS_CUTS_DEVICE=1 python -m pytest tests/test_gpu_points.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_points_cuts_device.log
