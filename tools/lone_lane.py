"""Developer tool (GPU box): per-step latency of the flight kernel as a function of how many trajectories are resident
(1 lane, 1 warp, 1 block per SM, ..., full) and of the kernel variant."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from erpl_monte_carlo_sim_b200 import _abi, _lib  # noqa: E402

eng = _lib.Engine(0)
md, blk, wind, _ = bench.make_workload("planar", 56832, 0)
eng.set_model(md)
variants = [dict(), dict(block_threads=128, blocks_per_sm=1, cold_state_in_smem=-1), dict(block_threads=128, blocks_per_sm=1, cold_state_in_smem=1),
            dict(block_threads=64, blocks_per_sm=1, cold_state_in_smem=-1), dict(block_threads=128, blocks_per_sm=3, cold_state_in_smem=1)]
sizes = [int(x) for x in os.environ.get("SIZES", "1,4736,18944,56832").split(",")]
for kw in variants:
    for n in sizes:
        best = None
        for rep in range(2):
            out, iout = eng.run_batch(np.ascontiguousarray(blk[:, :n]), np.ascontiguousarray(wind[:n]), opts=_lib.run_opts(**kw))
            c = eng.counters()
            if best is None or c["flight_ms"] < best["flight_ms"]:
                best = dict(c)
        ns = iout[_abi.IOUT["n_steps"]]; fn = iout[_abi.IOUT["first_nan_step"]]
        work = np.where(fn >= 0, np.minimum(fn, ns), ns)
        print(json.dumps({"cfg": kw, "n": n, "flight_ms": round(best["flight_ms"], 3), "max_work": int(work.max()),
                          "us_per_step_of_longest": round(best["flight_ms"] * 1e3 / work.max(), 3),
                          "gsteps_per_s": round(best["rk4_steps"] / best["flight_ms"] / 1e6, 3)}), flush=True)
