#!/bin/bash
# Round 2, session 2, validation on one GPU: all GPU tests, the full bench, launch list and one full ncu capture
set -x
mkdir -p gpurun_out
step() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAIL:-4} gpurun_out/$name.log | cut -c1-600; if [ $rc -ne 0 ]; then echo "STOP at $name"; exit 1; fi; }
step s2_gputests_final 900 python -m pytest tests/test_gpu_observations.py tests/test_gpu_points.py tests/test_gpu_sht.py tests/test_gpu_transformcl.py -x -q -m gpu
TAIL=1 step s2_bench_final 600 python bench.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02_bench_int8.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-chain --no-extra > gpurun_out/bench_under_ncu_int8.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sht_legendre_ozaki -c 1 -o gpurun_out/prof_oz4096 -f python tools/probe_ozaki.py 4096:8191 8 0 > gpurun_out/ncu_oz4096.log 2>&1; echo "ncu rc=$?"
