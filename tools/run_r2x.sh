for args in "--planar-samples 0 --c3-large 0 --c4-samples 300000" "--planar-samples 20000 --c3-large 0 --c4-samples 300000"; do
  echo "== $args"
  timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline $args > gpurun_out/r2x.json 2> gpurun_out/r2x.err; echo "rc=$?"; grep "EmcError" gpurun_out/r2x.err | tail -1 | cut -c1-600
  python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2x.json').read().strip().splitlines()[-1])
    for k,v in d.get('secondary',{}).items(): print(k, {kk:v[kk] for kk in v if kk in ('samples_per_gpu','trajectories_per_s','flight_ms_per_rank','wall_ms')})
except Exception as e: print('no json', e)
PY
done
