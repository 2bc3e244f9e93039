#!/bin/bash
# smoke() with the INT8 leg, lensing probe (INT8 refinement syntheses at nside 4096) and the lensing / full-size tests
set -x
mkdir -p gpurun_out
step() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAIL:-4} gpurun_out/$name.log | cut -c1-600; if [ $rc -ne 0 ]; then echo "STOP at $name"; exit 1; fi; }
step s2_smoke 120 python __graft_entry__.py smoke
step s2_lens_int8 200 python tools/probe_lensing.py 4096 3 4
GLB_LEGENDRE=fp64 step s2_lens_fp64 200 python tools/probe_lensing.py 4096 3 4
step s2_lens_tests 600 python -m pytest tests/test_gpu_lensing.py tests/test_gpu_fullsize.py tests/test_gpu_chain.py -x -q -m gpu
