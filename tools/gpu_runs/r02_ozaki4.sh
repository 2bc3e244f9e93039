#!/bin/bash
# INT8 Legendre path, careful bring-up: every step must pass before the next one runs
set -x
mkdir -p gpurun_out
step() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -4 gpurun_out/$name.log; if [ $rc -ne 0 ]; then echo "STOP at $name"; exit 1; fi; }
step oz_a 40 python tools/probe_ozaki.py 32:64,128:255 4,8 0
GLB_OZ_DEBUG=1 step oz_b 40 python tools/probe_ozaki.py 512:1023 8 0
GLB_OZ_DEBUG=1 step oz_c 60 python tools/probe_ozaki.py 1024:2047,2048:4095 8 0
step oz_big 120 python tools/probe_ozaki.py 2048:4095,4096:8191 4,8 2
step oz_tests 400 python -m pytest tests/test_gpu_int8.py -x -q -m gpu
if [ "$1" = "ncu" ]; then
step ncu_oz 300 ncu --set full --clock-control none --import-source on -k regex:sht_legendre_ozaki -c 1 -o gpurun_out/prof_oz -f python tools/probe_ozaki.py 2048:4095 8 0
fi
