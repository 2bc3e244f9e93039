#!/bin/bash
# Round 2, first call (2 GPUs): the fused Legendre + transpose over peer memory, written in round 1
# after the GPU budget was spent.  gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_runs/r02_p2p.sh'
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15 | tee gpurun_out/r02_p2p_tests.log
# all-to-all form against peer stores, same sizes (bit-identical check + timings in the output)
for mode in "" "--p2p"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    tests/dist/msplit_check.py $mode 256 1024 4096 2>&1 | grep -v "^W\|^\*" | tee -a gpurun_out/r02_p2p_timings.log
done
# memory checker on the small case (peer stores must stay inside the mapped blocks)
timeout 600 compute-sanitizer --tool memcheck python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
  --master-addr 127.0.0.1 --master-port 29543 tests/dist/msplit_check.py --p2p 8 48 2>&1 | tail -20 | tee gpurun_out/r02_p2p_memcheck.log
