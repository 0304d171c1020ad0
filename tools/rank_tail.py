"""Developer tool (GPU box): per-seed-range flight time and work distribution of the C3 workload — the per-rank batches of the
multi-GPU bench — to separate data-driven stragglers from hardware variation."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from erpl_monte_carlo_sim_b200 import _abi, _lib  # noqa: E402

n = int(os.environ.get("N", "100000"))
eng = _lib.Engine(0)
for r in range(int(os.environ.get("RANKS", "8"))):
    md, blk, wind, _ = bench.make_workload("c3", n, r * n)
    eng.set_model(md)
    best = 1e9; best_off = 1e9
    for rep in range(3):
        out, iout = eng.run_batch(blk, wind, opts=_lib.run_opts(lane_yield=False))
        best_off = min(best_off, eng.counters()["flight_ms"])
    for rep in range(3):
        out, iout = eng.run_batch(blk, wind)
        best = min(best, eng.counters()["flight_ms"])
    ns, fn = iout[_abi.IOUT["n_steps"]].astype(np.int64), iout[_abi.IOUT["first_nan_step"]].astype(np.int64)
    c = eng.counters()
    work = np.where(fn >= 0, np.minimum(fn, ns), ns)
    print(json.dumps({"rank": r, "flight_ms": round(best, 2), "flight_ms_without_hand_back": round(best_off, 2), "handed_back": int(eng.counters()["yielded"]), "rk4_steps": int(c["rk4_steps"]), "work_mean": float(work.mean()),
                      "work_p99": float(np.percentile(work, 99)), "work_max": int(work.max()),
                      "top5": np.sort(work)[-5:].tolist()}), flush=True)
