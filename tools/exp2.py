import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from erpl_monte_carlo_sim_b200 import _abi, _lib
import util
eng = _lib.Engine(0)
z = util.golden("mc_solid_csv"); md = _abi.model_from_npz(z); eng.set_model(md)
for n in (2048, 12500, 50000, 100000):
    sc, w = util.synth(z, n, seed=11)
    out, iout = eng.run_batch(sc, w)
    c = eng.counters()
    steps = iout[0]; fn = iout[3]
    real = np.where(fn >= 0, 0, steps)
    # real steps of NaN lanes = steps - replay; estimate via total
    print(json.dumps({"n": n, "flight_ms": round(c["flight_ms"], 2), "steps": c["rk4_steps"], "replay": c["replay_steps"],
                      "nan_frac": float((fn >= 0).mean()), "max_real_nonnan": int(real.max()),
                      "term": np.bincount(iout[1], minlength=5).tolist()}), flush=True)
    if n == 2048:
        ref = util.hostseam_batch(md, sc, w, nan_ff=True)
        print("iout equal vs seam:", [bool(np.array_equal(iout[k], ref[1][k])) for k in range(5)], "mismatch", int((iout[0] != ref[1][0]).sum()))
        # per-sample real steps on GPU for NaN samples: first_nan to all-nan gap unknown; print first_nan stats
        nanidx = np.flatnonzero(fn >= 0)
        print("first_nan min/mean/max", fn[nanidx].min(), fn[nanidx].mean(), fn[nanidx].max())
