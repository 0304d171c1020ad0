#!/bin/bash
# developer tool (GPU box): headline numbers of the current build: C3 100k, C3 1M, planar 100k
for args in "--samples-per-gpu 100000" "--samples-per-gpu 1000000" "--workload planar --samples-per-gpu 100000"; do
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras $args "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:12], d['config']['samples_per_gpu'], 'traj/s', round(d['value']), 'Gsteps/s', round(d['rk4_steps_per_s']/1e9,3))"
done
