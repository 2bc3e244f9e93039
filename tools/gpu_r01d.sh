set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/run_config.py 4 --shells 8 --lensing 2>gpurun_out/cfg4_2gpu.err | tee gpurun_out/cfg4_2gpu.json
tail -5 gpurun_out/cfg4_2gpu.err
