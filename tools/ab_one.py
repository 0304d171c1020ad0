"""Developer tool (GPU box): headline numbers of ONE build of libemc.so (EMC_LIB=... selects it): flight-kernel time on
C3 100 k / C3 1 M / planar 100 k, a lone trajectory's step latency, and a parity check against the reference goldens.
Workloads are cached in /tmp so that several builds can be compared in one gpurun call (tools/ab.sh)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import util  # noqa: E402
from erpl_monte_carlo_sim_b200 import _abi, _lib  # noqa: E402

CACHE = "/tmp/emc_ab_cache.npz"
if not os.path.isfile(CACHE):
    md, blk, wind, _ = bench.make_workload("c3", 100_000, 0)
    _, pblk, pwind, _ = bench.make_workload("planar", 100_000, 0)
    np.savez(CACHE, blk=blk, wind=wind, pblk=pblk, pwind=pwind)
z = np.load(CACHE)
md, _, _, _ = bench.make_workload("c3", 8, 0)
eng = _lib.Engine(0)
eng.set_model(md)
kw = json.loads(os.environ.get("EMC_AB_OPTS", "{}"))
opts = _lib.run_opts(**kw)
res = {"lib": os.path.basename(_lib.SO_PATH), "opts": kw}


def best_of(blk, wind, reps=3):
    b = None
    for _ in range(reps):
        eng.run_batch(blk, wind, opts=opts)
        c = eng.counters()
        if b is None or c["flight_ms"] < b["flight_ms"]:
            b = dict(c)
    return b


c = best_of(z["blk"], z["wind"], 4)
o100, io100 = eng.run_batch(z["blk"], z["wind"], opts=opts)
import hashlib  # noqa: E402
res["c3_100k_sha"] = hashlib.sha1(np.ascontiguousarray(o100).tobytes() + np.ascontiguousarray(io100).tobytes()).hexdigest()[:12]
del o100, io100
res["handovers"] = c.get("handovers"); res["yielded"] = c.get("yielded"); res["parked"] = c.get("parked"); res["strict_ms"] = round(c.get("strict_ms", 0.0), 3); res["strict_steps"] = c.get("strict_steps"); res["c3_100k_ms"] = round(c["flight_ms"], 3); res["c3_100k_gsteps"] = round(c["rk4_steps"] / c["flight_ms"] / 1e6, 3)
big_b = np.ascontiguousarray(np.tile(z["blk"], (1, 8))); big_w = np.ascontiguousarray(np.tile(z["wind"], (8, 1, 1)))
c = best_of(big_b, big_w, 2)
res["c3_800k_ms"] = round(c["flight_ms"], 3); res["c3_800k_gsteps"] = round(c["rk4_steps"] / c["flight_ms"] / 1e6, 3)
del big_b, big_w
c = best_of(z["pblk"], z["pwind"], 2)
res["planar_100k_ms"] = round(c["flight_ms"], 3); res["planar_100k_gsteps"] = round(c["rk4_steps"] / c["flight_ms"] / 1e6, 3)
out, iout = eng.run_batch(np.ascontiguousarray(z["pblk"][:, :1]), np.ascontiguousarray(z["pwind"][:1]), opts=opts)
c = eng.counters()
res["lone_us_per_step"] = round(c["flight_ms"] * 1e3 / max(int(iout[0, 0]), 1), 3)
# parity: the 64 golden C3 flights of the reference
g = util.golden("mc_solid_csv")
eng.set_model(_abi.model_from_npz(g))
o, io = eng.run_batch(g["scalars"], g["wind"], opts=opts)
res["golden_int_equal"] = bool(np.array_equal(io, g["iout"]))
try:
    res["golden_max_err"] = float(util.assert_summary_close(o, g["out"], what="ab"))
except AssertionError as e:
    res["golden_max_err"] = "FAIL: " + str(e)[:200]
print(json.dumps(res), flush=True)
