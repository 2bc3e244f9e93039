set -x
python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "spin2 or single" 2>&1 | tail -8
python -m pytest tests/test_gpu_sht.py tests/test_gpu_lensing.py -x -q -m gpu 2>&1 | tail -4
python tools/probe_lensing.py 2048 0
python tools/probe_lensing.py 4096 3
