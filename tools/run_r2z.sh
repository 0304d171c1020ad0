export EMC_DEBUG=1
timeout 120 python tools/c4_probe.py 200000 eng 2>&1 | grep "ok\|FAILED\|first\|emc\]" | cut -c1-400
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; grep "emc\]\|Error" gpurun_out/r2z_bench.err | sort | uniq -c | head
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['kernel_ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print({k:(v['trajectories_per_s']) for k,v in d['e2e_api'].items()})
for k,v in d.get('secondary',{}).items(): print(k, {kk:v[kk] for kk in v if kk in ('samples_per_gpu','trajectories_per_s','rk4_steps_per_s','flight_ms_per_rank','wall_ms','fp64_roofline_frac','hbm_write_GBps','waves_of_resident_lanes')})
PY
