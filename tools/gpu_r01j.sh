set -x
python -m pytest tests/test_gpu_fields.py -x -q -m gpu 2>&1 | tail -4
python tools/run_config.py 4 --shells 60 --ncorr 59 --no-galaxies 2>/dev/null | tee gpurun_out/cfg4_ncorr59.json
python tools/run_config.py 4 --shells 60 --ncorr 3 --no-galaxies 2>/dev/null | tee gpurun_out/cfg4_ncorr3.json
