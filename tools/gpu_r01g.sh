set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/bench_d.err | tee gpurun_out/bench_d.json
