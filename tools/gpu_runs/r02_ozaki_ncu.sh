#!/bin/bash
# ncu capture of the INT8 Legendre kernel (nside 2048, 8 maps)
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sht_legendre_ozaki -c 1 -o gpurun_out/prof_oz -f python tools/probe_ozaki.py 2048:4095 ${1:-8} 0 > gpurun_out/ncu_oz.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_oz.log
