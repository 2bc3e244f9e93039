#!/bin/bash
# Round-2 ncu evidence (one GPU):  gpurun --timeout 1500 -- 'bash tools/gpu_runs/prof_r02.sh'
set -x
mkdir -p gpurun_out
# launch list of the bench command (headline + e2e arm + other_stages; the chain is measured without a profiler)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-chain > gpurun_out/bench_under_ncu_r02.log 2>&1
# launch list of one stacked shear_from_convergence (4 planes, nside 4096, niter 3)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_lensing4096.csv python tools/probe_lensing.py 4096 3 4 > gpurun_out/lens_under_ncu_r02.log 2>&1
# full captures: batched spin kernel (4 planes), list-mode sampling kernels, long-ring FFT
ncu --set full --clock-control none --import-source on -k regex:spin_legendre_synth -s 4 -c 1 -o gpurun_out/prof_r02_spin4 -f python tools/probe_lensing.py 4096 0 4 > gpurun_out/ncu_r02_spin.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:points_fill_list|points_count_scan' -s 8 -c 3 -o gpurun_out/prof_r02_sampling -f python bench.py --steps 1 --warmup 0 --no-cpu --no-chain > gpurun_out/ncu_r02_sampling.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sht_ringfft_synth_long -c 1 -o gpurun_out/prof_r02_longfft -f python tools/probe_sht.py 8192:2047 1 > gpurun_out/ncu_r02_longfft.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
