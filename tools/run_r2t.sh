nvidia-smi -L | head -1
timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['kernel_ms_per_step'], 'e2e', d['e2e']['value'])
PY
