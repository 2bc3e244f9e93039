#!/bin/bash
# INT8 Legendre path: quick accuracy / timing / tests / short bench, stopping at the first failure
set -x
mkdir -p gpurun_out
step() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAIL:-4} gpurun_out/$name.log; if [ $rc -ne 0 ]; then echo "STOP at $name"; exit 1; fi; }
GLB_OZ_DEBUG=1 step oz_c 60 python tools/probe_ozaki.py 512:1023,2048:4095 8 0
step oz_big 120 python tools/probe_ozaki.py 4096:8191 4,8 2
step oz_tests 400 python -m pytest tests/test_gpu_int8.py -x -q -m gpu
TAIL=1 step bench_short 300 python bench.py --no-chain --no-cpu --no-extra
