#!/bin/bash
# INT8 Legendre path: accuracy + timing (optionally an ncu capture with "ncu" as first argument)
set -x
mkdir -p gpurun_out
timeout 120 python tools/probe_ozaki.py 32:64,128:255,512:1023 4,8 1 > gpurun_out/oz_small.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/oz_small.log
timeout 300 python tools/probe_ozaki.py 2048:4095,4096:8191 4,8 2 > gpurun_out/oz_big.log 2>&1; echo "rcbig=$?"; tail -4 gpurun_out/oz_big.log
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sht_legendre_ozaki -c 1 -o gpurun_out/prof_oz -f python tools/probe_ozaki.py 2048:4095 8 0 > gpurun_out/ncu_oz.log 2>&1; echo "rc=$?"
fi
