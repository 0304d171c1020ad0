#!/bin/bash
# developer tool: regenerate the text summaries under profiles/ from an .ncu-rep of the bench workload
# usage: tools/refresh_profiles.sh gpurun_out/prof.ncu-rep <lane RK4 steps in the launch> [prefix, default r2]
set -e
rep=$1; steps=${2:-216894169}; pre=${3:-r2}
python tools/ncu_summary.py "$rep" "$steps" > profiles/${pre}_flight_kernel_bench.txt 2>&1
ncu -i "$rep" --page raw --csv > profiles/${pre}_flight_kernel_bench_raw.csv 2>/dev/null
ncu -i "$rep" --page source --csv --print-source sass > /tmp/sass_page.csv 2>/dev/null
python - "$steps" "$pre" <<'PY' > profiles/${pre}_flight_kernel_opcodes.txt
import csv, collections, re, sys
steps = float(sys.argv[1])
rows = list(csv.reader(open('/tmp/sass_page.csv')))
h = rows[1]; iS = h.index("Source"); iE = h.index("Instructions Executed"); iN = h.index("# Samples")
ops = collections.Counter(); samp = collections.Counter(); static = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= iE: continue
    try: e = int(r[iE]); s = int(r[iN])
    except ValueError: continue
    t = re.sub(r'^@!?U?P\d+\s+', '', r[iS].strip())
    op = t.split()[0].split('.')[0] if t else '?'
    ops[op] += e; samp[op] += s; static[op] += 1; tot += e
lanes = None
for l in open(f'profiles/{sys.argv[2]}_flight_kernel_bench.txt'):
    if l.startswith('smsp__thread_inst_executed_per_inst_executed'): lanes = float(l.split()[1])
wd = steps / lanes * 4
print(f"emc_flight_kernel<128,3,2,1,1>, C3 100 k samples (ncu --set full source page, SASS view; same capture as {sys.argv[2]}_flight_kernel_bench.txt)")
print(f"warp instructions executed {tot}; per warp-level derivative evaluation {tot / wd:.0f}")
fp64 = sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"FP64-pipe instructions per derivative {fp64 / wd:.0f} ({100 * fp64 / tot:.1f} %): at 2 issue cycles each and 3 warps per scheduler the pipe alone needs {3 * 2 * fp64 / wd:.0f} cycles per derivative round")
print(f"{'opcode':10s} {'per deriv':>10s} {'share':>7s} {'stall samples':>14s} {'static':>7s}")
for k, v in ops.most_common(28):
    print(f"{k:10s} {v / wd:10.1f} {100 * v / tot:6.1f}% {100 * samp[k] / sum(samp.values()):13.1f}% {static[k]:7d}")
PY
