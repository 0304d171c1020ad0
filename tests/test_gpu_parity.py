"""GPU (-m gpu): the CUDA engine, called through the C ABI (libemc.so via ctypes), against
  * the golden vectors of the unmodified Python reference (tests/golden, oracle/make_golden.py),
  * the C oracle (oracle/emc_oracle.c) on larger seeded batches,
  * size-independent properties at full batch size (determinism, shard invariance, scheduling
    invariance).
Bar (BASELINE.json north_star): integers exact (step counts, termination codes, apogee indices),
FP64 summaries within 1e-6 relative."""
import numpy as np
import pytest

import oracle_lib as O
import util
from erpl_monte_carlo_sim_b200 import _abi, _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", util.DERIV_SETS)
def test_gpu_derivative(engine, name):
    z = util.golden(name)
    engine.set_model(_abi.model_from_npz(z))
    sd, ch = engine.derivative_debug(z["scalars"], z["wind"] if z["wind"].size else None, z["t"], z["state"], z["chute_in"])
    ref = z["state_dot"]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.nanmax(np.abs(sd - ref) / np.maximum(scale, 1e-300)) < 1e-12
    assert np.array_equal(np.isnan(sd), np.isnan(ref))
    assert np.array_equal(ch, z["chute_out"])


def test_gpu_single_flights(engine):
    z = util.golden("flights_single")
    for name in z["names"]:
        md, sc, wind, ref, iref = util.single_case(z, str(name))
        engine.set_model(md)
        out, iout = engine.run_batch(sc, wind)
        np.testing.assert_array_equal(iout, iref, err_msg=str(name))
        util.assert_summary_close(out, ref, what=str(name))


def test_gpu_series_on_the_reference_states(engine):
    """_extract_results (simulator.py:496-552) on the device, evaluated at the REFERENCE's own stored states of every golden
    single flight — all of them, the diverging tail included (the API test compares the engine's own states, which the
    blow-up separates from the reference's after ~2 500 steps; here nothing is amplified)."""
    z = util.golden("flights_single")
    for name in z["names"]:
        name = str(name)
        md, sc, wind, ref, iref = util.single_case(z, name)
        idx, rows, sref = util.series_reference(z, name)
        engine.set_model(md)
        got = engine.extract_series(sc, wind, rows.copy())
        assert got.shape == (_abi.SERIES_COUNT, idx.size)
        util.assert_series_close(got, sref, name)


@pytest.mark.parametrize("name", util.MC_SETS)
def test_gpu_mc_sets(engine, name):
    z = util.golden(name)
    engine.set_model(_abi.model_from_npz(z))
    out, iout = engine.run_batch(z["scalars"], z["wind"])
    np.testing.assert_array_equal(iout, z["iout"])
    util.assert_summary_close(out, z["out"], what=name)
    c = engine.counters()
    assert c["kernel_launches"] == 4 and c["refills"] == z["scalars"].shape[1]      # rail, flight, strict continuation (concurrent consumer + sweep)
    nan_ff = z["iout"][_abi.IOUT["first_nan_step"]] >= 0
    assert c["rk4_steps"] + c["strict_steps"] + c["replay_steps"] == int(z["iout"][0].sum())
    assert (c["replay_steps"] > 0) == bool(nan_ff.any())


def test_gpu_tape(engine):
    z = util.golden("flights_single")
    for name in ("c1b_example_liquid_csv", "c1c_planar_liquid"):
        md, sc, wind, ref, iref = util.single_case(z, name)
        engine.set_model(md)
        out, iout, tape = engine.run_tape(sc, wind)
        assert tape.shape[0] == iref[0, 0] + 1
        np.testing.assert_array_equal(iout, iref)
        util.assert_summary_close(out, ref, what=name)
        idx = z[name + "__series__idx"]
        ref_rows = z[name + "__series__tape"]
        ok = np.abs(ref_rows) < 1e15
        scale = np.maximum(np.maximum(np.abs(ref_rows), np.abs(ref_rows).max(axis=0, keepdims=True) * 1e-3), 1e-9)
        assert np.max((np.abs(tape[idx] - ref_rows) / scale)[ok]) < 1e-6


def test_gpu_tape_capacity_error(engine):
    z = util.golden("flights_single")
    md, sc, wind, ref, iref = util.single_case(z, "c1b_example_liquid_csv")
    engine.set_model(md)
    with pytest.raises(_lib.EmcError, match="EMC_ERR_CAPACITY"):
        engine.run_tape(sc, wind, cap=100)


def test_gpu_scheduling_invariance(engine):
    """Results do not depend on how lanes are scheduled.  Within one kernel binary (refill threshold,
    NaN fast-forward, grid size) outputs are bit-identical: each trajectory is integrated by one lane
    alone.  Other launch-bound variants are different compilations (different FMA contraction) and
    must agree to the parity tolerance with identical integer results."""
    z = util.golden("mc_liquid_default")
    engine.set_model(_abi.model_from_npz(z))
    base = engine.run_batch(z["scalars"], z["wind"])
    for kw in (dict(refill_threshold=32), dict(refill_threshold=8), dict(refill_threshold=2)):
        got = engine.run_batch(z["scalars"], z["wind"], opts=_lib.run_opts(**kw))
        np.testing.assert_array_equal(got[0], base[0], err_msg=str(kw))
        np.testing.assert_array_equal(got[1], base[1], err_msg=str(kw))
    for kw in (dict(block_threads=64), dict(block_threads=128, blocks_per_sm=3, cold_state_in_smem=False),
               dict(block_threads=128, blocks_per_sm=1), dict(block_threads=128, blocks_per_sm=4),
               dict(block_threads=128, cold_state_in_smem=False), dict(block_threads=128, blocks_per_sm=3, cold_state_in_smem=1)):
        got = engine.run_batch(z["scalars"], z["wind"], opts=_lib.run_opts(**kw))
        np.testing.assert_array_equal(got[1], base[1], err_msg=str(kw))
        util.assert_summary_close(got[0], base[0], what=str(kw))
    few = np.flatnonzero(z["iout"][_abi.IOUT["first_nan_step"]] >= 0)[:2]
    a = engine.run_batch(z["scalars"][:, few], z["wind"][few], opts=_lib.run_opts(nan_fast_forward=False))
    np.testing.assert_array_equal(a[0], base[0][:, few])
    np.testing.assert_array_equal(a[1], base[1][:, few])
    assert engine.counters()["replay_steps"] == 0


def test_gpu_tail_compaction_is_invisible(engine):
    """Once the work queue is empty, sparse warps hand their trajectories (lane records in shared memory) to one collector
    warp per block.  A batch of a few thousand flights of very different lengths (some NaN runs, some one-step flights):
    outputs with and without compaction are bit-identical, and the tape of a handed-over trajectory is complete."""
    z = util.golden("mc_liquid_default")
    engine.set_model(_abi.model_from_npz(z))
    rep = 40
    sc = np.ascontiguousarray(np.tile(z["scalars"], (1, rep))); wind = np.ascontiguousarray(np.tile(z["wind"], (rep, 1, 1)))
    sc[_abi.IN["q2"], ::7] = 0.0; sc[_abi.IN["q0"], ::7] = 1.0                 # horizontal launches: one-step flights in between
    on = engine.run_batch(sc, wind, opts=_lib.run_opts(compaction=True))
    assert engine.counters()["handovers"] > 50                                   # the tail really was compacted
    off = engine.run_batch(sc, wind)
    assert engine.counters()["handovers"] == 0
    np.testing.assert_array_equal(on[1], off[1]); np.testing.assert_array_equal(on[0], off[0])
    n = z["scalars"].shape[1]
    keep = np.arange(n) % 7 != 0
    np.testing.assert_array_equal(on[1][:, :n][:, keep], z["iout"][:, keep])
    longest = np.argsort(on[1][_abi.IOUT["n_steps"]] - np.maximum(on[1][_abi.IOUT["first_nan_step"]], 0))[-6:]
    engine.tape_request(longest, 50, 1300)
    engine.run_batch(sc, wind, opts=_lib.run_opts(compaction=True))
    rows, cnt = engine.tape_fetch()
    for k, i in enumerate(longest):
        if on[1][_abi.IOUT["first_nan_step"], i] < 0:
            ns = int(on[1][_abi.IOUT["n_steps"], i])
            assert cnt[k] == ns // 50 + 1 + (1 if ns % 50 else 0)
            assert rows[k, cnt[k] - 1, 0] == on[0][_abi.OUT["flight_time"], i] and rows[k, cnt[k] - 1, 3] == on[0][_abi.OUT["final_z"], i]


_synth = util.synth


@pytest.mark.parametrize("name,n", [("mc_solid_csv", 768), ("mc_liquid_default", 512), ("mc_planar_solid", 24)])
def test_gpu_vs_oracle_seeded_batch(engine, name, n):
    z = util.golden(name)
    md = _abi.model_from_npz(z)
    sc, wind = _synth(z, n, seed=2024)
    if name.startswith("mc_planar"):
        wind[:, :, 1] = 0.0
    engine.set_model(md)
    out, iout = engine.run_batch(sc, wind)
    ref, iref = O.batch(md, sc, wind)
    # step count, termination code, apogee index, first-NaN index, rail steps: exact on every sample (round 1 allowed 1 %
    # of blown-up flights to differ; the strict continuation closed that gap)
    np.testing.assert_array_equal(iout, iref)
    sens = util.oracle_sensitivity(md, sc, wind)
    assert np.mean(np.isinf(sens).any(axis=0)) < 0.02
    util.assert_summary_close(out, ref, what=name, sens=sens)            # NaN runs included, max |omega| too


def test_gpu_full_size_properties(engine):
    """BASELINE config C3 size (100k samples, Solid + 6-knot wind table): determinism, shard invariance,
    valid termination codes, counters consistent with the per-sample step counts."""
    z = util.golden("mc_solid_csv")
    md = _abi.model_from_npz(z)
    n = 100_000
    sc, wind = _synth(z, n, seed=7)
    engine.set_model(md)
    out, iout = engine.run_batch(sc, wind)
    c = engine.counters()
    steps = iout[_abi.IOUT["n_steps"]].astype(np.int64)
    assert c["rk4_steps"] + c["strict_steps"] + c["replay_steps"] == int(steps.sum())
    assert 0 < c["parked"] <= n and c["strict_steps"] >= c["parked"]          # blown-up flights finished by the strict continuation
    assert c["refills"] == n
    term = iout[_abi.IOUT["termination"]]
    assert np.all((term >= 1) & (term <= 4))
    assert np.all(steps >= 1)
    ft = out[_abi.OUT["flight_time"]]
    assert np.allclose(ft, steps * 0.005, rtol=0, atol=1e-6)          # t accumulates n_steps * dt
    assert np.all(out[_abi.OUT["rail_exit_speed"]] > 0)
    out2, iout2 = engine.run_batch(sc, wind, opts=_lib.run_opts(refill_threshold=16))
    assert np.array_equal(iout, iout2) and np.array_equal(out, out2, equal_nan=True)
    h = n // 2                                                          # two shards == one batch
    oa, ia = engine.run_batch(sc[:, :h].copy(), wind[:h])
    ob, ib = engine.run_batch(sc[:, h:].copy(), wind[h:])
    assert np.array_equal(np.concatenate([ia, ib], 1), iout)
    assert np.array_equal(np.concatenate([oa, ob], 1), out, equal_nan=True)
    # spot-check 256 of the 100k against the oracle
    pick = np.random.RandomState(1).choice(n, 256, replace=False)
    ref, iref = O.batch(md, sc[:, pick].copy(), wind[pick].copy())
    np.testing.assert_array_equal(iout[:, pick], iref)
    sens = util.oracle_sensitivity(md, sc[:, pick].copy(), wind[pick].copy())
    util.assert_summary_close(out[:, pick], ref, what="100k spot check", sens=sens)


def test_gpu_lane_hand_back_is_invisible(engine):
    """A batch larger than the resident lanes: after 900 stored states the flights of a warp whose attitude oscillation is
    growing put their state aside and their lanes start unstarted samples; the records are resumed once the queue is empty
    (emc_counters.yielded).  Outputs with and without the hand-back are bit-identical on the headline workload; launch ->
    landing flights (settled attitude) never hand back; a batch that fits the resident lanes has nothing to hand back for."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    md, blk, wind, _ = bench.make_workload("c3", 100_000, 0)
    engine.set_model(md)
    on = engine.run_batch(blk, wind)
    c = engine.counters()
    assert c["yielded"] > 20_000 and c["refills"] == 100_000
    steps = on[1][_abi.IOUT["n_steps"]].astype(np.int64)
    assert c["rk4_steps"] + c["strict_steps"] + c["replay_steps"] == int(steps.sum())      # a resumed flight repeats no step
    off = engine.run_batch(blk, wind, opts=_lib.run_opts(lane_yield=False))
    assert engine.counters()["yielded"] == 0
    np.testing.assert_array_equal(on[1], off[1]); np.testing.assert_array_equal(on[0], off[0])
    # the downsampled batch tape of flights that are handed back and resumed is the tape of the undisturbed flights
    picks = np.arange(0, 100_000, 1563, dtype=np.int64)
    tapes = []
    for kw in (dict(), dict(lane_yield=False)):
        engine.tape_request(picks, 50, 200)
        engine.run_batch(blk, wind, opts=_lib.run_opts(**kw))
        tapes.append((engine.counters()["yielded"], engine.counters()["tape_rows"]) + tuple(engine.tape_fetch()))
    assert tapes[0][0] > 20_000 and tapes[1][0] == 0 and tapes[0][1] == tapes[1][1] > 0
    np.testing.assert_array_equal(tapes[0][3], tapes[1][3])
    for k in range(len(picks)):
        m = int(tapes[0][3][k])
        np.testing.assert_array_equal(tapes[0][2][k, :m], tapes[1][2][k, :m])
    small = engine.run_batch(np.ascontiguousarray(blk[:, :40_000]), np.ascontiguousarray(wind[:40_000]))
    assert engine.counters()["yielded"] == 0
    np.testing.assert_array_equal(small[1], on[1][:, :40_000]); np.testing.assert_array_equal(small[0], on[0][:, :40_000])
    md, blk, wind, _ = bench.make_workload("planar", 60_000, 0)
    engine.set_model(md)
    engine.run_batch(blk, wind)
    assert engine.counters()["yielded"] == 0


def test_gpu_bench_workload_valid_flights_exact(engine):
    """On the headline workload itself (BASELINE C3, reference dispersions): the set of valid (non-outlier) flights is the
    oracle's, every valid flight has the oracle's integer outputs exactly and its summaries within 1e-6 (10x the oracle's
    own one-ulp sensitivity for the few valid flights that pass through a numerical blow-up), and EVERY flight — the
    blown-up ones included (the speed overflows to inf/NaN within one RK4 step; the strict continuation
    follows the reference's operation order there) — has the oracle's integer outputs."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from erpl_monte_carlo_sim_b200 import MonteCarloAnalyzer
    md, blk, wind, _ = bench.make_workload("c3", 6000, 300000)
    engine.set_model(md)
    out, iout = engine.run_batch(blk, wind)
    ref, iref = O.batch(md, blk, wind)
    OUT = _abi.OUT
    mask = lambda o: MonteCarloAnalyzer.outlier_mask(o[OUT["apogee_altitude"]], o[OUT["range"]], o[OUT["flight_time"]])
    bad_gpu, bad_ref = mask(out), mask(ref)
    np.testing.assert_array_equal(bad_gpu, bad_ref)
    valid = ~bad_ref
    assert valid.sum() >= 100
    np.testing.assert_array_equal(iout[:, valid], iref[:, valid])
    sens = util.oracle_sensitivity(md, blk, wind)
    util.assert_summary_close(out[:, valid], ref[:, valid], what="valid C3 flights", sens=sens[:, valid])
    well = valid & np.all(sens < 1e-8, axis=0)
    assert well.sum() >= 0.9 * valid.sum()
    util.assert_summary_close(out[:, well], ref[:, well], what="well-conditioned valid C3 flights")       # plain 1e-6
    # measured on 50 000 samples of this workload (tools/parity_sweep.py): no sample differs in an integer output
    np.testing.assert_array_equal(iout, iref)
    # every flight that stays finite (valid or not) within the conditioning-aware tolerance; the ones whose speed reaches
    # inf/NaN are covered above by the identical outlier sets (their diagnostics differ by inf-vs-NaN category only)
    tame = np.isfinite(out[OUT["max_speed"]]) & np.isfinite(ref[OUT["max_speed"]])
    assert tame.mean() > 0.5
    util.assert_summary_close(out[:, tame], ref[:, tame], what="finite C3 flights", sens=sens[:, tame])


def test_gpu_edge_cases(engine):
    z = util.golden("mc_solid_csv")
    md = _abi.model_from_npz(z)
    engine.set_model(md)
    out, iout = engine.run_batch(z["scalars"][:, :0].copy(), z["wind"][:0])          # empty batch
    assert out.shape == (_abi.OUT_COUNT, 0)
    out1, iout1 = engine.run_batch(z["scalars"][:, 5:6].copy(), z["wind"][5:6])      # ragged: one sample
    np.testing.assert_array_equal(iout1[:, 0], z["iout"][:, 5])
    out33, iout33 = engine.run_batch(z["scalars"][:, :33].copy(), z["wind"][:33])    # one lane past a warp
    np.testing.assert_array_equal(iout33, z["iout"][:, :33])
    # shared wind table (stride 0) == replicated table
    w0 = z["wind"][3]
    a = engine.run_batch(z["scalars"][:, :8].copy(), w0, wind_shared=True)
    b = engine.run_batch(z["scalars"][:, :8].copy(), np.repeat(w0[None], 8, 0))
    np.testing.assert_array_equal(a[0], b[0])
    # max_time shorter than the rail time: zero RK4 steps, termination = max_time
    md2 = dict(md); md2["max_time"] = 0.5
    engine.set_model(md2)
    o, i = engine.run_batch(z["scalars"][:, :4].copy(), z["wind"][:4])
    ro, ri = O.batch(md2, z["scalars"][:, :4].copy(), z["wind"][:4])
    np.testing.assert_array_equal(i, ri)
    assert np.all(i[_abi.IOUT["n_steps"]] == 0) and np.all(i[_abi.IOUT["termination"]] == 4)
    util.assert_summary_close(o, ro, what="zero-step flights")


def test_gpu_errors(engine):
    e2 = _lib.Engine(0)
    z = util.golden("mc_readme_literal")
    with pytest.raises(_lib.EmcError, match="EMC_ERR_NO_MODEL"):
        e2.has_wind = True; e2.n_wind = 100
        e2.run_batch(z["scalars"], z["wind"])
    md = _abi.model_from_npz(z)
    bad = dict(md); bad["cd_mach"] = np.array(md["cd_mach"])[::-1].copy()
    with pytest.raises(_lib.EmcError, match="strictly increasing"):
        e2.set_model(bad)
    e2.close()
    with pytest.raises(_lib.EmcError, match="out of range"):
        _lib.Engine(99)


def test_gpu_math_helpers(engine):
    """The derivative kernel's reciprocal / rsqrt / sqrt (MUFU seed + Newton) and atan2 (minimax
    polynomial): accuracy in ulp against NumPy over the operand ranges the path produces."""
    rng = np.random.RandomState(0)
    x = np.concatenate([10.0 ** rng.uniform(-8, 12, 200000), rng.uniform(100.0, 200.0, 50000), [1.0, 4.0, 6.371e6, 1e-300, 1e300]])

    def ulps(got, ref):
        return np.abs(got - ref) / np.spacing(np.abs(ref))
    assert ulps(engine.math_debug(0, x), 1.0 / x).max() <= 1.5
    assert ulps(engine.math_debug(1, x), 1.0 / np.sqrt(x)).max() <= 2.0
    assert ulps(engine.math_debug(3, x), np.sqrt(x)).max() <= 2.0
    assert engine.math_debug(3, np.array([0.0]))[0] == 0.0 and np.isnan(engine.math_debug(3, np.array([np.nan]))[0])
    xe = np.concatenate([rng.uniform(-20, 2, 300000), rng.uniform(-700, 700, 20000), [0.0, -0.0, 1e-300, -745.0, 710.0]])
    got, ref = engine.math_debug(4, xe), np.exp(xe)
    ok = (xe > -700) & (xe < 700)
    assert ulps(got[ok], ref[ok]).max() <= 1.5
    assert got[-2] == 0.0 and np.isinf(got[-1]) and np.isnan(engine.math_debug(4, np.array([np.nan]))[0])
    xl = np.concatenate([rng.uniform(0.5, 2.0, 300000), 10.0 ** rng.uniform(-300, 300, 20000), [1.0, 0.75, 1.0551]])
    got, ref = engine.math_debug(5, xl), np.log(xl)
    err = np.abs(got - ref) / np.maximum(np.spacing(np.abs(ref)), 2.3e-17)         # near x = 1 log -> 0: absolute floor
    assert err.max() <= 4.0, err.max()
    assert got[-3] == 0.0 and np.isnan(engine.math_debug(5, np.array([-1.0]))[0]) and np.isinf(engine.math_debug(5, np.array([0.0]))[0])
    ang = rng.uniform(-np.pi, np.pi, 400000)
    r = 10.0 ** rng.uniform(-6, 6, ang.size)
    yy, xx = r * np.sin(ang), r * np.cos(ang)
    yy[:1000] = rng.normal(0, 1e-9, 1000) * xx[:1000]                       # small angles
    got, ref = engine.math_debug(2, xx, yy), np.arctan2(yy, xx)
    assert ulps(got, ref).max() <= 3.5, ulps(got, ref).max()
    sp_y = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 0.0, 1.0, -1.0, np.nan])
    sp_x = np.array([1.0, 1.0, -1.0, -1.0, 0.0, 0.0, 0.0, 1.0, -1.0, 1.0])
    got, ref = engine.math_debug(2, sp_x, sp_y), np.arctan2(sp_y, sp_x)
    np.testing.assert_allclose(got, ref, rtol=3e-16, atol=0, equal_nan=True)
    assert np.array_equal(np.signbit(got[:4]), np.signbit(ref[:4]))


def test_gpu_components_match_reference(engine):
    """Atmosphere, gravity, mass properties, aerodynamic coefficients and thrust evaluated by the device functions
    against the reference's own values (components.npz) — SURVEY §4 component KATs."""
    z = util.golden("components")
    models = {"liquid": _abi.model_from_npz(z, "liquid_"), "solid": _abi.model_from_npz(z, "solid_")}

    def ev(kind, comp, cols):
        engine.set_model(models[kind])
        return engine.component(comp, *cols)
    util.check_components(ev, z)


def test_gpu_irregular_samples_fly_the_strict_continuation(engine):
    """Samples the fast path does not accept (propellant mass negative, dry mass zero / negative / NaN) are handed to the
    strict continuation before their first step: their outputs are the oracle's, integers exactly."""
    z = util.golden("mc_solid_csv")
    md = _abi.model_from_npz(z)
    sc, wind = z["scalars"][:, :24].copy(), z["wind"][:24].copy()
    IN = _abi.IN
    sc[IN["prop_mass"], 1] = -3.0           # mass < dry_mass fires (simulator.py:315-318)
    sc[IN["prop_mass"], 5] = -1e-3
    sc[IN["dry_mass"], 9] = -50.0           # negative total mass: the reference flies on
    sc[IN["dry_mass"], 12] = 0.0
    sc[IN["prop_mass"], 12] = 0.0           # zero mass: division by zero in the first derivative
    sc[IN["dry_mass"], 17] = np.nan
    engine.set_model(md)
    out, iout = engine.run_batch(sc, wind)
    c = engine.counters()
    ref, iref = O.batch(md, sc, wind)
    np.testing.assert_array_equal(iout, iref)
    # summaries: the untouched samples under the usual rule; the six irregular ones (all of them end in the reference's
    # blow-up, where CUDA's libm against glibc's — 1-2 ulp per call — is amplified without bound: 6.49e36 came out 1.8e-4
    # apart) by category where non-finite and to 1 % on the finite values
    odd = np.zeros(sc.shape[1], bool); odd[[1, 5, 9, 12, 17]] = True
    sens = util.oracle_sensitivity(md, sc[:, ~odd], wind[~odd])
    util.assert_summary_close(out[:, ~odd], ref[:, ~odd], what="regular samples vs oracle", sens=sens)
    err = util.summary_errors(out[:, odd], ref[:, odd])        # NaN / inf / beyond-1e150 values compare by category (tests/util.py)
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err[worst] <= 1e-2, (f"irregular sample {np.flatnonzero(odd)[worst[1]]}, field {_abi.OUT_FIELDS[worst[0]]}: "
                                f"{out[:, odd][worst]!r} vs {ref[:, odd][worst]!r}")
    assert c["parked"] >= 5
