#!/bin/bash
# generate() with the look-ahead batch on a side stream: parity tests, then the short bench
set -x
mkdir -p gpurun_out
step() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAIL:-4} gpurun_out/$name.log | cut -c1-700; if [ $rc -ne 0 ]; then echo "STOP at $name"; exit 1; fi; }
step s2_pipe_tests 600 python -m pytest tests/test_gpu_fields.py -x -q -m gpu
TAIL=1 step bench_short 300 python bench.py --no-chain --no-cpu --no-extra
