"""Developer tool (GPU box, torchrun): the stream-ordered multi-GPU statistics chain against the pass-by-pass path and
against NumPy on the gathered samples.  torchrun --nproc-per-node N tools/dist_stats_check.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from erpl_monte_carlo_sim_b200 import _abi, stats as S  # noqa: E402
from erpl_monte_carlo_sim_b200.simulator import get_engine  # noqa: E402
from stats_numpy_backend import NumpyBackend  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
O = _abi.OUT
n = 40000 + 1000 * rank
rng = np.random.RandomState(100 + rank)
out = np.zeros((_abi.OUT_COUNT, n))
out[O["apogee_altitude"]] = rng.normal(9000 + 500 * rank, 3000, n); out[O["range"]] = np.abs(rng.normal(40000, 15000, n))
out[O["flight_time"]] = np.round(rng.normal(11, 1, n) * 200) / 200; out[O["final_x"]] = rng.normal(0, 2e4, n); out[O["final_y"]] = rng.normal(0, 4e4, n)
out[O["apogee_altitude"], ::7] = np.nan
eng = get_engine(local)
eng.upload_outputs(out)
a = S.device_statistics(eng, n, fused=True)
b = S.device_statistics(eng, n, fused=False)
t0 = time.perf_counter()
for _ in range(20):
    S.device_statistics(eng, n, fused=True)
torch.cuda.synchronize(); tf = (time.perf_counter() - t0) / 20
t0 = time.perf_counter()
for _ in range(20):
    S.device_statistics(eng, n, fused=False)
torch.cuda.synchronize(); tp = (time.perf_counter() - t0) / 20
cols = [torch.zeros(5, 41000 + 1000 * world, dtype=torch.float64, device="cuda") for _ in range(world)]
mine = torch.full((5, 41000 + 1000 * world), float("nan"), dtype=torch.float64, device="cuda")
mine[:, :n] = torch.from_numpy(out[[O["apogee_altitude"], O["range"], O["flight_time"], O["final_x"], O["final_y"]]]).cuda()
sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device="cuda"))
dist.all_gather(cols, mine)
allv = np.concatenate([c[:, :int(s.item())].cpu().numpy() for c, s in zip(cols, sizes)], axis=1)
ref = S.compute_statistics(NumpyBackend(*allv))
ok = True
for key in ("apogee_altitude", "range", "flight_time"):
    ok &= a[key]["percentiles"] == b[key]["percentiles"] == ref[key]["percentiles"]
    ok &= a[key]["min"] == ref[key]["min"] and a[key]["max"] == ref[key]["max"]
    ok &= abs(a[key]["mean"] - ref[key]["mean"]) <= 1e-12 * abs(ref[key]["mean"]) and abs(a[key]["std"] - ref[key]["std"]) <= 1e-11 * abs(ref[key]["std"])
ok &= a["n_samples"] == b["n_samples"] == ref["n_samples"] and a["n_total"] == ref["n_total"]
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print({"world": world, "all_ranks_ok": bool(flag.item()), "n_total": a["n_total"], "n_samples": a["n_samples"],
           "ms_stream_ordered_chain": round(tf * 1e3, 3), "ms_pass_by_pass": round(tp * 1e3, 3)}, flush=True)
dist.destroy_process_group()
