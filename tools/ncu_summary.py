"""Developer tool: condensed text summary of an .ncu-rep (raw metrics + per-source-line instruction shares)."""
import collections, csv, subprocess, sys

rep = sys.argv[1]
nsteps = float(sys.argv[2]) if len(sys.argv) > 2 else None      # lane RK4 steps in the profiled launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum"]
d = dict(zip(hdr, zip(vals, units)))
for k in want:
    if k in d:
        print(f"{k:85s} {d[k][0]:>22s} {d[k][1]}")
for k in hdr:
    if "issue_stalled" in k and "per_issue_active" in k and float(d[k][0] or 0) > 0.02:
        print(f"{k:85s} {d[k][0]:>22s}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; h = None
by = collections.Counter(); samp = collections.Counter(); text = {}; ops = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; h = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": h = r; iE = r.index("Instructions Executed"); iS = r.index("# Samples"); continue
    if h is None or r[0] == "": continue
    try: e = int(r[iE]); s = int(r[iS])
    except ValueError: continue
    key = (cur, int(r[0])); by[key] += e; samp[key] += s; text[key] = r[1].strip()[:100]
tot = sum(by.values()); ts = sum(samp.values()) or 1
inst = float(d["smsp__inst_executed.sum"][0])
print(f"\nsource-attributed instructions {tot} (kernel total {inst:.0f})")
if nsteps:
    lanes = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"][0])
    wd = nsteps / lanes * 4
    print(f"warp-level derivative evaluations ~ {wd:.3e}; warp instructions per derivative ~ {inst / wd:.0f}")
for k, v in by.most_common(45):
    per = f"{v / wd:7.1f}/deriv" if nsteps else ""
    print(f"{k[0][:18]:18s}{k[1]:5d} {per} {100 * v / tot:5.1f}%  stall-samples {100 * samp[k] / ts:5.1f}% | {text[k]}")
