"""Developer experiment (GPU box): throughput of the flight kernel on tiled golden sets."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from erpl_monte_carlo_sim_b200 import _abi, _lib
import util

def tile(z, keep, reps):
    sc = np.ascontiguousarray(np.tile(z["scalars"][:, keep], (1, reps)))
    w = np.ascontiguousarray(np.tile(z["wind"][keep], (reps, 1, 1)))
    return sc, w

def run(eng, label, sc, w, **kw):
    best = None
    for rep in range(2):
        out, iout = eng.run_batch(sc, w, opts=_lib.run_opts(**kw))
        c = eng.counters()
        if best is None or c["flight_ms"] < best["flight_ms"]:
            best = c
    n = sc.shape[1]
    print(json.dumps({"label": label, "n": n, "cfg": kw, "flight_ms": round(best["flight_ms"], 2), "rail_ms": round(best["rail_ms"], 3),
                      "steps": best["rk4_steps"], "replay": best["replay_steps"],
                      "gsteps_per_s": round(best["rk4_steps"] / best["flight_ms"] / 1e6, 4),
                      "traj_per_s": round(n / best["flight_ms"] * 1e3)}), flush=True)

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    eng = _lib.Engine(0)
    z = util.golden("mc_liquid_default"); md = _abi.model_from_npz(z)
    eng.set_model(md)
    ok = np.flatnonzero(z["iout"][3] < 0)
    allk = np.arange(64)
    if which in ("all", "A"):
        sc, w = tile(z, ok, 1024)
        for kw in (dict(), dict(refill_threshold=32), dict(block_threads=128, blocks_per_sm=3), dict(block_threads=128, blocks_per_sm=4), dict(block_threads=256), dict(block_threads=64)):
            run(eng, "A liquid100 no-NaN x1024", sc, w, **kw)
        sc, w = tile(z, ok, 4096)
        run(eng, "A liquid100 no-NaN x4096", sc, w)
    if which in ("all", "B"):
        sc, w = tile(z, allk, 1024)
        run(eng, "B liquid100 all x1024 ff", sc, w)
        sc, w = tile(z, allk, 128)
        run(eng, "C liquid100 all x128 noff", sc, w, nan_fast_forward=False)
        run(eng, "C liquid100 all x128 ff", sc, w)
    if which in ("all", "D"):
        z = util.golden("mc_planar_solid"); md = _abi.model_from_npz(z); eng.set_model(md)
        okp = np.flatnonzero(z["iout"][3] < 0)
        sc, w = tile(z, okp, 2048)
        run(eng, "D planar solid x2048", sc, w)
        run(eng, "D planar solid x2048", sc, w, refill_threshold=32)
        run(eng, "D planar solid x2048", sc, w, block_threads=128, blocks_per_sm=3)
        sc, w = tile(z, okp, 8192)
        run(eng, "D planar solid x8192", sc, w)
    if which == "ncu":
        sc, w = tile(z, ok, 1024)
        run(eng, "ncu A x1024", sc, w)
    if which == "ncu3":
        sc, w = tile(z, ok, 1024)
        run(eng, "ncu A x1024 (128,3)", sc, w, block_threads=128, blocks_per_sm=3)

if __name__ == "__main__":
    main()
