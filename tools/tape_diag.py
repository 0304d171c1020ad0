"""developer tool (GPU box): where does a one-flight run differ from the same sample inside a batch?
usage: PYTHONPATH=$PWD python tools/tape_diag.py"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import util
from erpl_monte_carlo_sim_b200 import _abi, _lib

eng = _lib.Engine(0)
z = util.golden("mc_solid_csv")
eng.set_model(_abi.model_from_npz(z))
sc, wind = z["scalars"], z["wind"]
out, iout = eng.run_batch(sc, wind)
print("batch counters", eng.counters())
eq = lambda a, b: bool(np.array_equal(a, b, equal_nan=True))
for nff in (True, False):
    o2, io2 = eng.run_batch(sc, wind, opts=_lib.run_opts(nan_fast_forward=nff, refill_threshold=1))
    print("batch again nan_ff", nff, "thr 1: equal", eq(out, o2), eq(iout[[0, 1, 2, 4]], io2[[0, 1, 2, 4]]))
for i in (0, 3, 17, 63):
    for nff in (True, False):
        o1, io1 = eng.run_batch(sc[:, i:i + 1], wind[i:i + 1], opts=_lib.run_opts(nan_fast_forward=nff))
        c = eng.counters()
        print(i, "n=1 batch nan_ff", nff, "equal", eq(out[:, i], o1[:, 0]), "parked", c["parked"], "strict_steps", c["strict_steps"], "rk4", c["rk4_steps"])
    o1, io1, full = eng.run_tape(sc[:, i:i + 1], wind[i])
    c = eng.counters()
    print(i, "tape equal", eq(out[:, i], o1[:, 0]), "parked", c["parked"], "strict_steps", c["strict_steps"], "rk4", c["rk4_steps"], "n_steps", int(io1[0, 0]))
    # pairs: the sample with one neighbour
    for j in (1, 2, 9):
        sel = [i, j]
        o1, io1 = eng.run_batch(sc[:, sel], wind[sel])
        print(i, "pair with", j, "equal", eq(out[:, i], o1[:, 0]), eq(out[:, j], o1[:, 1]), "n_steps", int(iout[0, j]), "first_nan", int(iout[3, j]))

print("---- rows of the batch tape against the one-flight tape")
picks = np.array([0, 3, 5, 17, 63], np.int64)
for stride in (7,):
    cap = 60000 // stride + 4
    eng.tape_request(picks, stride, cap)
    out, iout = eng.run_batch(sc, wind)
    rows, cnt = eng.tape_fetch()
    for k, i in enumerate(picks):
        o1, io1, full = eng.run_tape(sc[:, i:i + 1], wind[i])
        t_rail = o1[_abi.OUT["rail_exit_time"], 0]
        m = int(cnt[k])
        strided = np.arange(0, m - 1) * stride
        got = rows[k, :m - 1]
        want = np.column_stack([full[strided, 0] - t_rail, full[strided, 1:4]])
        bad = np.flatnonzero(~np.all((got == want) | ((got != got) & (want != want)), axis=1))
        print(dict(sample=int(i), n_steps=int(iout[0, i]), rows=m, n_bad=len(bad), first_bad_step=int(bad[0]) * stride if len(bad) else None,
                   out_equal=eq(out[:, i], o1[:, 0]), final_row_equal=eq(full[-1, 1:4], out[[_abi.OUT["final_x"], _abi.OUT["final_y"], _abi.OUT["final_z"]], i])))
