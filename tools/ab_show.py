"""print the A/B lines of a gpurun log as a table.  usage: python tools/ab_show.py gpurun_out/x.log"""
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(f"{d['lib']:24s} {json.dumps(d['opts']):28s} c3_100k {d['c3_100k_ms']:7.3f} ms  c3_800k {d['c3_800k_gsteps']:6.3f} G  planar {d['planar_100k_gsteps']:6.3f} G  lone {d['lone_us_per_step']:6.3f} us  int_eq {d['golden_int_equal']}  err {d['golden_max_err'] if isinstance(d['golden_max_err'], str) else format(d['golden_max_err'], '.2e')}")
