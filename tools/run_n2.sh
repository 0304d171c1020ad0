#!/bin/bash
# developer tool (GPU box with 2 GPUs): the 2-GPU tests, then bench.py at N = 2 under torchrun -> gpurun_out/<tag>_bench_n2.json
tag=${1:-n2}
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_tape_and_shards.py -m gpu -q -k "device_group or sharded" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "N=2 rc=$?"; tail -c 300 gpurun_out/${tag}_bench_n2.err
