set -x
python -m pytest tests/test_gpu_sht.py tests/test_gpu_lensing.py tests/test_gpu_fields.py -x -q -m gpu 2>&1 | tail -5
python tools/probe_analysis.py 2048
GLB_AN_R=4 python tools/probe_analysis.py 2048
python tools/probe_analysis.py 4096
GLB_AN_R=4 python tools/probe_analysis.py 4096
python bench.py --steps 3 --warmup 3 --no-cpu --no-extra 2>gpurun_out/bench_c.err | tee gpurun_out/bench_c.json
ncu --set full --clock-control none --import-source on -k regex:legendre_analysis -s 1 -c 1 -o gpurun_out/prof_r01c_analysis -f python tools/probe_analysis.py 2048 > gpurun_out/ncu_r01c_an.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spin_legendre -s 1 -c 1 -o gpurun_out/prof_r01c_spin -f python tools/probe_lensing.py 2048 0 > gpurun_out/ncu_r01c_spin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sht_ringfft_synth -s 3 -c 1 -o gpurun_out/prof_r01c_ringfft -f python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_r01c_fft.log 2>&1
ls -la gpurun_out | tail -5
