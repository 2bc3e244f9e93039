#!/bin/bash
# INT8 Legendre path integrated: GPU tests (new file first), then the 1-GPU bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int8.py -x -q -m gpu > gpurun_out/s2_int8_tests.log 2>&1; echo "int8 tests rc=$?"; tail -15 gpurun_out/s2_int8_tests.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_int8.py > gpurun_out/s2_gputests2.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s2_gputests2.log
timeout 600 python bench.py > gpurun_out/s2_bench_int8.json 2> gpurun_out/s2_bench_int8.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/s2_bench_int8.err; head -c 2500 gpurun_out/s2_bench_int8.json
