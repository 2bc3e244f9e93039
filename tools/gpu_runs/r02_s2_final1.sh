#!/bin/bash
# Round 2, session 2, validation on one GPU: all GPU tests, the full bench, launch list and one full ncu capture
set -x
mkdir -p gpurun_out
step() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAIL:-4} gpurun_out/$name.log | cut -c1-600; if [ $rc -ne 0 ]; then echo "STOP at $name"; exit 1; fi; }
step s2_gputests_final 900 python -m pytest tests -x -q -m gpu
TAIL=1 step s2_bench_final 600 python bench.py
