set -x
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_sht.py tests/test_gpu_lensing.py -x -q -m gpu 2>&1 | tail -15
python tools/probe_analysis.py 2048
python tools/probe_analysis.py 4096
python tools/probe_lensing.py 4096 3
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_lens4096_c.csv python tools/probe_lensing.py 4096 3 > gpurun_out/lens4096_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:galaxy_shear -c 1 -o gpurun_out/prof_r01c_gshear -f python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_r01c_gshear.log 2>&1
