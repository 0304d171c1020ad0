nvidia-smi -L | head -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --planar-samples 0 --c3-large 0 > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"; tail -c 500 gpurun_out/r3a_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3a_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['kernel_ms_per_step'], 'e2e', d['e2e']['value'])
print(json.dumps(d['e2e_api'])[:900])
for k,v in d.get('secondary',{}).items(): print(k, {kk:v[kk] for kk in v if kk in ('samples_per_gpu','trajectories_per_s','rk4_steps_per_s','flight_ms_per_rank','wall_ms')})
PY
