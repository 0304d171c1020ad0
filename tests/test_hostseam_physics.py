"""CPU: the engine's own device physics (csrc/emc_physics.cuh) compiled by g++ through the test seam
(tests/hostseam) against the reference goldens and the oracle.  This is the same code the CUDA kernels
inline; it lets a container without a GPU catch parity regressions before GPU time is spent.  The
seam is not part of the product (libemc.so has no CPU path)."""
import numpy as np
import pytest

import oracle_lib as O
import util
from erpl_monte_carlo_sim_b200 import _abi


@pytest.mark.parametrize("name", util.DERIV_SETS)
def test_seam_derivative(name):
    z = util.golden(name)
    md = _abi.model_from_npz(z)
    sd, ch = util.hostseam_derivative(md, z["scalars"], z["wind"] if z["wind"].size else None, z["t"], z["state"], z["chute_in"])
    ref = z["state_dot"]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.nanmax(np.abs(sd - ref) / np.maximum(scale, 1e-300)) < 1e-13
    assert np.array_equal(np.isnan(sd), np.isnan(ref))
    assert np.array_equal(ch, z["chute_out"])


def test_seam_single_flights():
    z = util.golden("flights_single")
    for name in z["names"]:
        md, sc, wind, ref, iref = util.single_case(z, str(name))
        out, iout = util.hostseam_batch(md, sc, wind)
        np.testing.assert_array_equal(iout, iref, err_msg=str(name))
        util.assert_summary_close(out, ref, what=str(name))


@pytest.mark.parametrize("name", util.MC_SETS)
def test_seam_mc_sets(name):
    z = util.golden(name)
    md = _abi.model_from_npz(z)
    out, iout = util.hostseam_batch(md, z["scalars"], z["wind"])
    np.testing.assert_array_equal(iout, z["iout"])
    util.assert_summary_close(out, z["out"], what=name)


def test_seam_nan_fast_forward_is_exact():
    """Integrating the all-NaN tail (reference behaviour) and replaying only `t` give identical outputs."""
    z = util.golden("mc_liquid_default")
    md = _abi.model_from_npz(z)
    nan_runs = np.flatnonzero(z["iout"][_abi.IOUT["first_nan_step"]] >= 0)[:3]
    assert nan_runs.size > 0
    sc, w = z["scalars"][:, nan_runs], z["wind"][nan_runs]
    a = util.hostseam_batch(md, sc, w, nan_ff=True)
    b = util.hostseam_batch(md, sc, w, nan_ff=False)
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[0], b[0])


def test_seam_ballistic_nan_fast_forward():
    """Mode 2 of the fast-forward (altitude NaN after burnout: x, y coast ballistically): same step
    counts and, to rounding, the same outputs as grinding through the derivative like the reference.
    Flown the way the engine flies it: fast path up to the blow-up trigger, then the strict continuation, which is where
    the pattern arises (an infinite body-x force through a rotation matrix with an exactly zero entry gives F_z = NaN
    and F_y = -inf; the fast path's quaternion rotation turns the same state all-NaN at once, which is mode 1)."""
    z = util.golden("mc_solid_csv")
    md = _abi.model_from_npz(z)
    sc, wind = util.synth(z, 2048, seed=11)
    sc, wind = np.ascontiguousarray(sc[:, 640:768]), np.ascontiguousarray(wind[640:768])     # holds sample 724: vy = -inf, chute out
    full = util.hostseam_batch_strict(md, sc, wind, nan_ff=False)
    fast = util.hostseam_batch_strict(md, sc, wind, nan_ff=True)
    ref = O.batch(md, sc, wind)
    np.testing.assert_array_equal(full[1], fast[1])
    np.testing.assert_array_equal(full[1], ref[1])          # with the strict continuation every integer output is the oracle's
    assert fast[2].max() <= 8 < full[2].max()               # strict steps: a handful when the NaN tail is replayed in closed form
    OUT = _abi.OUT
    nanz = np.isnan(full[0][OUT["final_z"]])
    ballistic = nanz & ~np.isnan(full[0][OUT["final_y"]])
    assert ballistic.sum() >= 1, "the seeded batch must contain an altitude-NaN flight that keeps coasting in x/y"
    rows = [i for k, i in OUT.items() if k != "max_abs_omega"]
    util.assert_summary_close(fast[0][rows], full[0][rows], rtol=1e-9, what="fast-forward vs full integration")
    sens = util.oracle_sensitivity(md, sc, wind)         # blown-up synthetic flights amplify one ulp beyond 1e-6
    util.assert_summary_close(fast[0], np.where(np.isin(np.arange(ref[0].shape[0]), rows)[:, None], ref[0], fast[0]),
                              what="fast-forward vs oracle", sens=sens)
    # max|omega| keeps its value at the fast-forward point: never above the full integration's
    assert np.all(fast[0][OUT["max_abs_omega"]] <= full[0][OUT["max_abs_omega"]] * (1 + 1e-12))


def test_closed_form_time_replay_is_bit_exact():
    """The NaN fast-forward advances `t += dt` binade by binade in integer arithmetic; it must land on
    exactly the t, step count and burnout time that ~57 k floating-point additions produce."""
    import ctypes as C
    HS = util.hostseam_lib()
    HS.hs_replay.argtypes = [C.c_double] * 5 + [C.c_int, C.POINTER(C.c_double)]
    rng = np.random.RandomState(3)
    cases = []
    for rail_steps in (0, 1, 60, 67, 87, 94, 120):
        t_rail = 0.0
        for _ in range(rail_steps):
            t_rail += 0.01
        for k in (0, 1, 1868, 2371, 2914, 30000):
            t0 = t_rail
            for _ in range(k):
                t0 += 0.005
            for burn in (13.7, 14.906103286384978, 15.63, 0.0, 400.0, float("nan")):
                cases.append((t0, t_rail, 0.005, 300.0, burn))
    cases += [(0.3 + rng.rand(), 0.3, dt, mt, 14.0 + rng.rand()) for dt in (0.005, 0.0025, 0.001953125, 0.00390625, 0.0037, 1e-3)
              for mt in (300.0, 299.9975, 17.3, 1.0, 0.2)]
    cases += [(2.0 ** -1030, 0.0, 0.005, 1.0, 0.5), (299.999, 0.87, 0.005, 300.0, 14.9), (300.0, 0.87, 0.005, 300.0, 14.9)]
    a = (C.c_double * 5)(); b = (C.c_double * 5)()
    for c in cases:
        HS.hs_replay(*c, 1, a); HS.hs_replay(*c, 0, b)
        assert list(a)[:2] == list(b)[:2] and a[4] == b[4], (c, list(a), list(b))
        assert a[3] == b[3] and (a[2] == b[2] or a[3] == 0.0), (c, list(a), list(b))


def test_seam_series_match_reference():
    z = util.golden("flights_single")
    for name in z["names"]:
        name = str(name)
        md, sc, wind, ref, iref = util.single_case(z, name)
        idx, rows, sref = util.series_reference(z, name)
        got = util.hostseam_series(md, sc, wind, rows.copy())
        util.assert_series_close(got, sref, name)


def test_seam_components_match_reference():
    z = util.golden("components")
    models = {"liquid": _abi.model_from_npz(z, "liquid_"), "solid": _abi.model_from_npz(z, "solid_")}
    util.check_components(lambda kind, comp, cols: util.hostseam_component(models[kind], comp, cols), z)


def test_troposphere_pressure_series_matches_the_power_law():
    """The engine evaluates the layered pressure law as polynomials in altitude (DevModel.tp_c below the tropopause, the
    at_* segments above it up to 100 km); they must agree with the reference's expressions (environment.py:26-103, through
    the oracle) to rounding over every layer and at every layer edge, hand over to the exp/log path outside without a
    jump, and switch themselves off for an atmosphere they cannot represent."""
    import ctypes as C
    z = util.golden("components")
    md = _abi.model_from_npz(z, "liquid_")
    edges = np.array([11000.0, 20000.0, 25000.0, 32000.0, 32000.0 + 48.65 / 0.0028, 100000.0])
    alt = np.concatenate([np.linspace(-2500.0, 12000.0, 3001), [-2000.0, -2000.0000001, 11000.0, 11000.0000001, 0.0, -0.0],
                          np.linspace(11000.0, 105000.0, 9401), edges, np.nextafter(edges, 0), np.nextafter(edges, 1e9),
                          [1e6, np.inf]])
    got = util.hostseam_component(md, 0, (alt,))
    L = O.lib()
    m, keep = _abi.pack_model(md)
    ref = np.empty((3, alt.size))
    T, p, rho = C.c_double(), C.c_double(), C.c_double()
    for i, a in enumerate(alt):
        L.orc_atmosphere(C.byref(m), C.c_double(a), C.byref(T), C.byref(p), C.byref(rho))
        ref[:, i] = (T.value, p.value, rho.value)
    np.testing.assert_allclose(got[:3], ref, rtol=3e-15)
    assert np.isnan(util.hostseam_component(md, 0, (np.array([np.nan]),))[:3]).all()
    # a lapse rate the series cannot cover in 17 terms falls back to exp/log and still matches
    md2 = dict(md); md2["temperature_lapse_rate"] = 0.02
    got2 = util.hostseam_component(md2, 0, (alt[:3001:50],))
    m2, keep2 = _abi.pack_model(md2)
    for i, a in enumerate(alt[:3001:50]):
        L.orc_atmosphere(C.byref(m2), C.c_double(a), C.byref(T), C.byref(p), C.byref(rho))
        if np.isfinite(p.value):
            np.testing.assert_allclose(got2[1, i], p.value, rtol=2e-14)


def test_mach_union_grid_reproduces_both_tables():
    """Cd(M) and CP(M) are looked up on the union of their knot vectors with one search; values must be the ones
    np.interp gives on the separate tables (rocket.py:105-108,156-157), for shared, interleaved and disjoint knots,
    at the knots themselves and beyond both ends."""
    import ctypes as C
    z = util.golden("components")
    base = _abi.model_from_npz(z, "liquid_")
    rng = np.random.RandomState(5)
    L = O.lib()
    L.orc_aero_coefficients.restype = None
    for trial in range(12):
        n_cd, n_cp = rng.randint(2, 17), rng.randint(2, 17)
        cdm = np.sort(rng.choice(np.arange(0, 60), n_cd, replace=False)) * 0.1
        cpm = np.sort(rng.choice(np.arange(0, 60), n_cp, replace=False)) * 0.1 + (0.0 if trial % 3 else 0.05)
        if trial == 4:
            cpm = cdm[-1] + 1.0 + np.arange(n_cp) * 0.5        # disjoint ranges
        md = dict(base)
        md["cd_mach"], md["cd0"], md["cda"] = cdm, 0.2 + rng.rand(n_cd), rng.rand(n_cd) * 5
        md["cp_mach"], md["cp_shift"] = cpm, rng.randn(n_cp) * 0.2
        mach = np.concatenate([cdm, cpm, rng.rand(200) * 8.0 - 0.5, np.nextafter(cdm, 0), np.nextafter(cpm, 100)])
        mach = np.abs(mach)
        alpha = rng.randn(mach.size) * 0.1; beta = rng.randn(mach.size) * 0.1
        got = util.hostseam_component(md, 2, (mach, alpha, beta, 1.3, 1.0, 1.0))
        m, keep = _abi.pack_model(md)
        c = (C.c_double * 6)()
        ref = np.empty((6, mach.size))
        for i in range(mach.size):
            L.orc_aero_coefficients(C.byref(m), C.c_double(mach[i]), C.c_double(alpha[i]), C.c_double(beta[i]), C.c_double(1.3),
                                    C.c_int(1), C.c_double(1.0), c)
            ref[:, i] = list(c)
        np.testing.assert_allclose(got[:6], ref, rtol=2e-14, atol=1e-16, err_msg=f"trial {trial}")
        np.testing.assert_allclose(got[5], ref[5], rtol=4e-16)            # CP: table value + cp_location, one FMA rounding apart


@pytest.mark.parametrize("name", ["mc_solid_csv", "mc_liquid_default", "mc_planar_solid", "mc_planar_liquid", "mc_readme_literal"])
def test_strict_continuation_reproduces_the_oracle_bit_for_bit(name):
    """csrc/emc_strict.cuh (the code emc_strict_kernel runs for trajectories whose blow-up is under way) compiled by g++:
    flown from the oracle's rail-exit state it must give the oracle's flight outputs BIT FOR BIT — integers, summaries and
    the running maxima/minima, NaN patterns included.  This pins its operation order to the reference's."""
    z = util.golden(name)
    md = _abi.model_from_npz(z)
    n = min(z["scalars"].shape[1], 4 if "planar" in name else 64)
    sc, wind = z["scalars"][:, :n], z["wind"][:n]
    ref, iref = O.batch(md, sc, wind)
    out, iout, steps = util.hostseam_batch_strict(md, sc, wind, out_with_rail=ref)
    np.testing.assert_array_equal(iout[:4], iref[:4])
    flight = [i for i, k in enumerate(_abi.OUT_FIELDS) if not k.startswith("rail_") and not k.startswith("wind_at")]
    np.testing.assert_array_equal(out[flight], ref[flight])                    # NaN == NaN positionally
    np.testing.assert_array_equal(iout[:4], z["iout"][:4, :n])                 # ... and the reference's own integers
    assert steps.sum() > 0


def test_fast_path_hands_blown_up_flights_to_the_strict_continuation():
    """The pair the GPU runs (fast kernel until |v| > 1e7 m/s or |omega| > 1000 rad/s, then the strict code), on 3 000 seeded
    C3 samples the goldens do not contain: every integer output equals the oracle's (the fast path alone leaves ~0.6 %
    of them different, all blown-up flights whose overflow pattern depends on the operation order); flights that land
    are never parked."""
    z = util.golden("mc_solid_csv")
    md = _abi.model_from_npz(z)
    sc, wind = util.synth(z, 3000, 11)
    ref, iref = O.batch(md, sc, wind)
    fast, ifast = util.hostseam_batch(md, sc, wind)
    both, iboth, steps = util.hostseam_batch_strict(md, sc, wind)
    assert int((ifast[:4] != iref[:4]).any(axis=0).sum()) >= 5               # the defect this exists for
    np.testing.assert_array_equal(iboth[:4], iref[:4])
    assert 0.4 < (steps > 0).mean() <= 1.0 and steps.max() <= 6                 # only the last steps of a blow-up are strict
    zp = util.golden("mc_planar_solid")
    out, iout, steps_p = util.hostseam_batch_strict(_abi.model_from_npz(zp), zp["scalars"][:, :4], zp["wind"][:4])
    landed = zp["iout"][_abi.IOUT["termination"], :4] == 1
    assert np.all(steps_p[landed & (zp["iout"][_abi.IOUT["first_nan_step"], :4] < 0) & (np.abs(zp["out"][_abi.OUT["max_speed"], :4]) < 1e4)] == 0)
    np.testing.assert_array_equal(iout, zp["iout"][:, :4])
