import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle"); sys.path.insert(0, ".")
import numpy as np, util
from erpl_monte_carlo_sim_b200 import _abi, _lib
import oracle_lib as O
eng = _lib.Engine(0)
for name, n in (("mc_solid_csv", 768), ("mc_liquid_default", 512)):
    z = util.golden(name); md = _abi.model_from_npz(z)
    sc, wind = util.synth(z, n, seed=2024)
    eng.set_model(md)
    out, iout = eng.run_batch(sc, wind)
    ref, iref = O.batch(md, sc, wind)
    nan_run = iref[_abi.IOUT["first_nan_step"]] >= 0
    i = _abi.OUT["max_abs_omega"]
    a, b = out[i, nan_run], ref[i, nan_run]
    same_cat = (np.isnan(a) == np.isnan(b)) & (np.isinf(a) == np.isinf(b))
    fin = np.isfinite(a) & np.isfinite(b)
    rel = np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)
    print(name, "nan runs", int(nan_run.sum()), "same category", int(same_cat.sum()), "finite both", int(fin.sum()),
          "bit equal", int((a == b).sum() + (np.isnan(a) & np.isnan(b)).sum()), "max rel (finite)", rel.max() if rel.size else None,
          "rel > 1e-6:", int((rel > 1e-6).sum()), "log10 |ref| range", np.log10(np.abs(b[fin])).min() if fin.any() else None, np.log10(np.abs(b[fin])).max() if fin.any() else None)
