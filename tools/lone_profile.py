"""Developer tool (GPU box): run ONE planar trajectory (for an ncu capture of the lone-warp instruction stream)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from erpl_monte_carlo_sim_b200 import _lib  # noqa: E402

eng = _lib.Engine(0)
md, blk, wind, _ = bench.make_workload(os.environ.get("WORKLOAD", "c3"), 64, int(os.environ.get("SEED0", "0")))
eng.set_model(md)
k = int(os.environ.get("PICK", "0"))
out, iout = eng.run_batch(np.ascontiguousarray(blk[:, k:k + 1]), np.ascontiguousarray(wind[k:k + 1]))
print(eng.counters(), iout[:, 0])
