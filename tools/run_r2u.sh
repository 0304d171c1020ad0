nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_tape_and_shards.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2u_bench2.json 2> gpurun_out/r2u_bench2.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2u_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['kernel_ms_per_step'], d['e2e']['value'])
print(json.dumps(d.get('e2e_api'))[:600])
for k,v in d.get('secondary',{}).items(): print(k, {kk:v[kk] for kk in v if kk in ('samples_per_gpu','trajectories_per_s','rk4_steps_per_s','flight_ms_per_rank','wall_ms','fp64_roofline_frac','hbm_write_GBps')})
PY
