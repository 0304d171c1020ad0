"""Developer tool (GPU box): one-off wide parity check of the CUDA path against the C oracle on the bench workloads
(more samples than the test-suite affords).  Prints one JSON line per workload."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import oracle_lib as O  # noqa: E402
import util  # noqa: E402
from erpl_monte_carlo_sim_b200 import _abi, _lib  # noqa: E402

eng = _lib.Engine(0)


def c4_workload(n, first_index):
    """BASELINE config C4 as the bench flies it: LiquidMotor, default dispersions, 100-knot stochastic wind per sample, drawn
    and perturbed on the device (Philox); the staged inputs are read back so that the oracle flies the same samples."""
    from erpl_monte_carlo_sim_b200 import LiquidMotor, MonteCarloAnalyzer, Rocket, StandardAtmosphere, WindModel, marshal
    mc = MonteCarloAnalyzer(Rocket(), LiquidMotor(), StandardAtmosphere(), WindModel())
    md = marshal.model_dict(mc.rocket, mc.motor, mc.atmosphere, mc._model_simulator(), mc._altitude_grid())
    eng.set_model(md)
    eng.generate_inputs(mc.dispersion_struct(bench.IC_C4), mc.philox_seed, first_index, n)
    sc, w = eng.staged_inputs(n, want_wind=True)
    return md, sc, w


for workload, n, seed0 in (("c3", int(os.environ.get("N_C3", "20000")), 300000), ("planar", int(os.environ.get("N_PLANAR", "1500")), 700000),
                           ("c4", int(os.environ.get("N_C4", "10000")), 5000000)):
    if n <= 0:
        continue
    if workload == "c4":
        md, blk, wind = c4_workload(n, seed0)
    else:
        md, blk, wind, _ = bench.make_workload(workload, n, seed0)
    eng.set_model(md)
    out, iout = eng.run_batch(blk, wind)
    t0 = time.time()
    ref, iref = O.batch(md, blk, wind)
    t_or = time.time() - t0
    same = np.all(iout == iref, axis=0)
    sens = util.oracle_sensitivity(md, blk, wind)
    err = util.summary_errors(out, ref)
    finite = np.isfinite(err)
    tol = np.maximum(1e-6, 10.0 * sens)
    bad = finite & (err > tol) & same[None, :]
    well = np.all(sens < 1e-9, axis=0) & same        # well-conditioned flights: plain 1e-6 rule, and how close we really are
    from erpl_monte_carlo_sim_b200 import MonteCarloAnalyzer
    OUT = _abi.OUT
    mask = lambda o: MonteCarloAnalyzer.outlier_mask(o[OUT["apogee_altitude"]], o[OUT["range"]], o[OUT["flight_time"]])
    mg, mr = mask(out), mask(ref)
    differ = ~same
    if os.environ.get("DUMP_DIFFER"):
        idx = np.flatnonzero(differ)
        np.savez(os.environ["DUMP_DIFFER"] + "_" + workload + ".npz", idx=idx, scalars=blk[:, idx], wind=wind[idx], out=out[:, idx], iout=iout[:, idx],
                 ref=ref[:, idx], iref=iref[:, idx])
    if os.environ.get("VERBOSE"):
        names = list(OUT)
        for f, i in list(zip(*np.nonzero(bad)))[:40]:
            print("  violation", int(i), names[f], "err %.2e sens %.2e" % (err[f, i], sens[f, i]), "ref", ref[f, i], "got", out[f, i],
                  "valid" if not mr[i] else "outlier", "steps", int(iref[0, i]), "first_nan", int(iref[3, i]), flush=True)
        ev = np.where(np.isfinite(err), err, 0.0)[:, ~mr]
        worst = np.argsort(-ev.max(axis=0))[:5]
        vi = np.flatnonzero(~mr)
        for w in worst:
            f = int(np.argmax(ev[:, w])); i = int(vi[w])
            print("  valid-worst", i, names[f], "err %.2e sens %.2e" % (err[f, i], sens[f, i]), "ref", ref[f, i], "got", out[f, i], flush=True)
        for i in np.flatnonzero(differ & ~(np.isnan(out[OUT["max_speed"]]) & np.isnan(ref[OUT["max_speed"]])))[:10]:
            print("  differ-not-both-nan", int(i), iout[:, i].tolist(), iref[:, i].tolist(), "max_speed", out[OUT["max_speed"], i], ref[OUT["max_speed"], i], flush=True)
    print(json.dumps({"workload": workload, "n": n, "first_seed": seed0, "integer_outputs_identical": int(same.sum()),
                      "integer_mismatch": int((~same).sum()),
                      "integer_mismatch_all_blown_up_in_both (max_speed inf or NaN)": bool(not np.any(np.isfinite(out[OUT["max_speed"]][differ])) and not np.any(np.isfinite(ref[OUT["max_speed"]][differ]))),
                      "valid_flights": int((~mr).sum()), "valid_sets_identical": bool(np.array_equal(mg, mr)),
                      "valid_flights_integer_exact": bool(np.array_equal(iout[:, ~mr], iref[:, ~mr])),
                      "valid_flights_max_scaled_error": float(np.nanmax(err[:, ~mr])) if (~mr).any() else None,
                      "violations_beyond_conditioning_aware_tolerance": int(bad.sum()),
                      "well_conditioned_flights": int(well.sum()),
                      "max_scaled_error_well_conditioned": float(np.nanmax(np.where(np.isfinite(err[:, well]), err[:, well], 0.0))) if well.any() else None,
                      "oracle_seconds": round(t_or, 1)}), flush=True)
