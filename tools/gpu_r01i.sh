set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/run_config.py 4 --lensing 2>gpurun_out/cfg4_8gpu.err | tee gpurun_out/cfg4_8gpu.json
tail -3 gpurun_out/cfg4_8gpu.err
