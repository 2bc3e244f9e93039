#!/bin/bash
# INT8 Legendre path, first runs: small sizes with a sanitizer, then accuracy + timing up to nside 4096
set -x
mkdir -p gpurun_out
timeout 120 python tools/probe_ozaki.py 32:64 4 0 > gpurun_out/oz_32.log 2>&1; echo "rc32=$?"; tail -5 gpurun_out/oz_32.log
timeout 300 compute-sanitizer --tool memcheck python tools/probe_ozaki.py 32:64 4 0 > gpurun_out/oz_32_memcheck.log 2>&1; echo "rcmc=$?"; tail -15 gpurun_out/oz_32_memcheck.log
timeout 120 python tools/probe_ozaki.py 128:255,512:1023 4,8 1 > gpurun_out/oz_mid.log 2>&1; echo "rcmid=$?"; tail -6 gpurun_out/oz_mid.log
timeout 300 python tools/probe_ozaki.py 2048:4095,4096:8191 4,8 2 > gpurun_out/oz_big.log 2>&1; echo "rcbig=$?"; tail -6 gpurun_out/oz_big.log
