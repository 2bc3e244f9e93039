set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_r01.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sht_legendre_synth -s 1 -c 1 -o gpurun_out/prof_r01_legendre -f python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_r01_leg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sht_ringfft_synth -s 3 -c 3 -o gpurun_out/prof_r01_ringfft -f python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_r01_fft.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:points_|multiplane|galaxy_shear' -c 9 -o gpurun_out/prof_r01_sampling -f python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_r01_sampling.log 2>&1
ls -la gpurun_out
