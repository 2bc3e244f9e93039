#!/bin/bash
set -x
mkdir -p gpurun_out
GLB_OZ_DEBUG=1 timeout 60 python tools/probe_ozaki.py 2048:4095 ${1:-8} 0 > gpurun_out/oz_dbg.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/oz_dbg.log
