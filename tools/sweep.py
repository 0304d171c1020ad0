"""Tuning sweep (developer tool, GPU box): times the flight kernel over launch configurations on
synthetic batches derived from the golden sets.  Prints one JSON line per configuration."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from erpl_monte_carlo_sim_b200 import _abi, _lib  # noqa: E402
import util  # noqa: E402
from test_gpu_parity import _synth  # noqa: E402


def main():
    n = int(os.environ.get("SWEEP_N", "100000"))
    sets = os.environ.get("SWEEP_SETS", "mc_solid_csv,mc_planar_solid").split(",")
    eng = _lib.Engine(0)
    tf, ms = eng.fp64_peak()
    print(json.dumps({"fp64_peak_tflops": tf, "ms": ms}), flush=True)
    for name in sets:
        z = util.golden(name)
        md = _abi.model_from_npz(z)
        nn = n if "planar" not in name else max(1024, n // 8)
        sc, wind = _synth(z, nn, seed=11)
        if "planar" in name:
            wind[:, :, 1] = 0.0
        eng.set_model(md)
        configs = [dict(), dict(refill_threshold=8), dict(refill_threshold=16), dict(refill_threshold=32),
                   dict(block_threads=64), dict(block_threads=256), dict(block_threads=128, blocks_per_sm=3),
                   dict(block_threads=128, blocks_per_sm=4), dict(block_threads=128, blocks_per_sm=3, refill_threshold=16),
                   dict(block_threads=128, blocks_per_sm=4, refill_threshold=16)]
        for kw in configs:
            best = None
            for rep in range(2):
                t0 = time.time()
                out, iout = eng.run_batch(sc, wind, opts=_lib.run_opts(**kw))
                wall = time.time() - t0
                c = eng.counters()
                if best is None or c["flight_ms"] < best["flight_ms"]:
                    best = dict(c); best["wall_ms"] = wall * 1e3
            steps = best["rk4_steps"]
            print(json.dumps({"set": name, "n": nn, "cfg": kw, "flight_ms": round(best["flight_ms"], 3),
                              "rail_ms": round(best["rail_ms"], 3), "wall_ms": round(best["wall_ms"], 2),
                              "steps": steps, "replay": best["replay_steps"],
                              "gsteps_per_s": round(steps / best["flight_ms"] / 1e6, 4),
                              "traj_per_s": round(nn / (best["flight_ms"] + best["rail_ms"]) * 1e3, 1),
                              "tflops_1600": round(steps * 1600 / best["flight_ms"] / 1e9, 3)}), flush=True)


if __name__ == "__main__":
    main()
