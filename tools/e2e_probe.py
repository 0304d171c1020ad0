"""Developer probe (GPU box): where the end-to-end step spends its time."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from erpl_monte_carlo_sim_b200 import _abi, _lib
md, blk, wind, _ = bench.make_workload("c3", 100000, 0)
eng = _lib.Engine(0); eng.set_model(md)
n = blk.shape[1]
pb, pw = torch.from_numpy(blk).pin_memory().numpy(), torch.from_numpy(wind).pin_memory().numpy()
po = torch.empty((_abi.OUT_COUNT, n), dtype=torch.float64).pin_memory().numpy(); pi = torch.empty((_abi.IOUT_COUNT, n), dtype=torch.int32).pin_memory().numpy()
for label, kw in (("pinned", dict(outputs=(po, pi))), ("pinned-in pageable-out", {})):
    for _ in range(2): eng.run_batch(pb, pw, **kw)
    t0 = time.perf_counter()
    for _ in range(5): eng.run_batch(pb, pw, **kw)
    dt = (time.perf_counter() - t0) / 5; c = eng.counters()
    print(label, "e2e ms", round(dt * 1e3, 2), "kernels ms", round(c["flight_ms"] + c["rail_ms"], 2))
for _ in range(2): eng.run_batch(blk, wind)
t0 = time.perf_counter()
for _ in range(5): eng.run_batch(blk, wind)
print("pageable e2e ms", round((time.perf_counter() - t0) / 5 * 1e3, 2))
