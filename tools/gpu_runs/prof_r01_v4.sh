set -x
# final round-1 evidence, sized to stay under the 64 MiB that travel back: launch list of the
# bench command + one full capture of the dominant kernel; the analysis / spin / ring-FFT /
# sampling captures of this round are prof_r01c_*.ncu-rep and prof_r01_sampling.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01_v4_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_r01_v4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sht_legendre_synth -s 1 -c 1 -o gpurun_out/prof_r01_v4_legendre -f python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_r01_v4_leg.log 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_r01_final.err | tee gpurun_out/bench_r01_final.json
python bench.py --impl reference --steps 1 --warmup 0 2>gpurun_out/bench_r01_ref.err | tee gpurun_out/bench_r01_ref.json
du -sh gpurun_out
