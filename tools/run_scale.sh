#!/bin/bash
# developer tool (GPU box with 8 GPUs): the 2-GPU tests, then bench.py at N = 8, 4, 2 under torchrun -> gpurun_out/<tag>_bench_n<N>.json
tag=${1:-scale}
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_tape_and_shards.py -m gpu -q -k "device_group or sharded" 2>&1 | tail -3
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
    bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
  echo "N=$n rc=$?"; tail -c 300 gpurun_out/${tag}_bench_n$n.err
done
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for n in (8, 4, 2):
    try:
        d = json.loads(open(f'gpurun_out/{tag}_bench_n{n}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(n, 'no line', e); continue
    print(n, {k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['kernel_ms_per_step'].get('flight_per_rank'))
    for k, v in d.get('secondary', {}).items():
        print('   ', k, {kk: v[kk] for kk in v if kk in ('samples_per_gpu', 'trajectories_per_s', 'rk4_steps_per_s', 'flight_ms_per_rank', 'wall_ms', 'fp64_roofline_frac')})
PY
