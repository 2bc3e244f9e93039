set -x
python -m pytest tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -15
python -m pytest tests/test_gpu_sht.py tests/test_gpu_lensing.py tests/test_gpu_fields.py tests/test_gpu_points.py -x -q -m gpu 2>&1 | tail -5
python tools/probe_sht.py 4096:8191 4
