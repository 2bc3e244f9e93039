#!/bin/bash
# the north-star chain on one GPU with a reduced number of shells first (look-ahead generate + lensing on one plan)
set -x
mkdir -p gpurun_out
step() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAIL:-3} gpurun_out/$name.log | cut -c1-700; if [ $rc -ne 0 ]; then echo "STOP at $name"; exit 1; fi; }
step s2_chain20 120 python tools/run_config.py 4 --lensing --shells 20
step s2_chain60 200 python tools/run_config.py 4 --lensing
