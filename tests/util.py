"""Shared helpers of the test-suite: golden loading and the parity comparison."""
import ctypes as C
import os
import subprocess

import numpy as np

from erpl_monte_carlo_sim_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
OUT = _abi.OUT
RTOL = 1e-6          # BASELINE.json north_star: 1e-6 relative in FP64

# vector-valued outputs are compared against the vector's norm; angles/times get a small floor
VECTORS = [("rail_exit_x", "rail_exit_y", "rail_exit_z"), ("rail_exit_vx", "rail_exit_vy", "rail_exit_vz"),
           ("wind_at_exit_u", "wind_at_exit_v", "wind_at_exit_w"), ("final_x", "final_y", "final_z"),
           ("final_vx", "final_vy", "final_vz")]
ANGLES = ["rail_exit_roll", "rail_exit_pitch", "rail_exit_yaw", "rail_exit_aoa", "rail_exit_sideslip", "max_abs_aoa"]
ANGLE_ATOL = 1e-9    # rad
OVERFLOW_REGIME = 1e150   # ~sqrt(DBL_MAX)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def summary_errors(out, ref):
    """Scaled error per (field, sample): |a-b| / scale, where scale is |ref| (scalars) or the norm of
    the reference vector (vector groups).  NaN-vs-NaN and exact equality (incl. inf) count as 0."""
    out = np.asarray(out, float); ref = np.asarray(ref, float)
    with np.errstate(all="ignore"):
        scale = np.abs(ref).copy()
        for grp in VECTORS:
            idx = [OUT[g] for g in grp]
            nrm = np.sqrt(np.sum(ref[idx] ** 2, axis=0))
            scale[idx] = nrm
        err = np.abs(out - ref) / np.maximum(scale, 1e-300)
        for a in ANGLES:
            i = OUT[a]
            err[i] = np.abs(out[i] - ref[i]) / np.maximum(scale[i], ANGLE_ATOL / RTOL)
        # time-like fields that are exactly 0 in the reference (apogee_time, burnout_time)
        for tname in ("apogee_time", "burnout_time", "rail_exit_time"):
            i = OUT[tname]
            err[i] = np.abs(out[i] - ref[i]) / np.maximum(scale[i], 1e-3)
    same = (out == ref) | (np.isnan(out) & np.isnan(ref))
    err[same] = 0.0
    err[np.isnan(out) != np.isnan(ref)] = np.inf
    err[np.isnan(err)] = np.inf
    # Overflow regime: a reference value beyond sqrt(DBL_MAX) is one multiplication away from inf, where a
    # fused multiply-add (single rounding, no intermediate overflow) legitimately turns inf/NaN into a
    # finite number or back.  Such values (the reference's super-exponential blow-ups) are
    # compared by category: the engine's value must be non-finite or beyond the same bound as well.
    with np.errstate(invalid="ignore"):
        over_ref = ~np.isfinite(ref) | (np.abs(ref) > OVERFLOW_REGIME)
        over_out = ~np.isfinite(out) | (np.abs(out) > OVERFLOW_REGIME)
    err[over_ref & over_out] = 0.0
    return err


def assert_summary_close(out, ref, rtol=RTOL, what="", sens=None):
    """Every field of every sample within rtol (scaled as in summary_errors).  `sens` (optional, from
    oracle_sensitivity) widens the bar on ILL-CONDITIONED flights only: where a 1-ulp change of one input
    moves the reference algorithm's own output by s > rtol/10, the bar for that field is 10*s."""
    err = summary_errors(out, ref)
    tol = np.full(err.shape, rtol)
    if sens is not None:
        tol = np.maximum(tol, 10.0 * sens)
    excess = np.where(err <= tol, 0.0, err / tol)
    worst = np.unravel_index(np.argmax(excess), excess.shape)
    assert excess[worst] == 0.0, (f"{what}: field {_abi.OUT_FIELDS[worst[0]]} sample {worst[1]}: "
                                  f"{out[worst]!r} vs {ref[worst]!r} (scaled err {err[worst]:.3e} > {tol[worst]:.3e})")
    return float(err.max())


def oracle_sensitivity(md, scalars, wind):
    """Self-conditioning of the reference algorithm, measured with the C oracle: the largest scaled change
    of every output field when ONE input (dry mass, attitude q0, thrust) moves by one ulp.  Well-conditioned
    flights give ~1e-13; the reference's super-exponentially diverging flights (F6/F7)
    give up to 1e-5 and beyond.  Samples whose step count itself flips get +inf (nothing can be asserted
    beyond the category)."""
    import oracle_lib as O
    o0, i0 = O.batch(md, scalars, wind)
    worst = np.zeros_like(o0)
    for field in ("dry_mass", "q0", "thrust_a"):
        sc1 = np.array(scalars, dtype=np.float64, copy=True)
        i = _abi.IN[field]
        sc1[i] = np.nextafter(sc1[i], np.inf)
        o1, i1 = O.batch(md, sc1, wind)
        s = summary_errors(o1, o0)
        s[:, np.any(i0 != i1, axis=0)] = np.inf
        worst = np.maximum(worst, s)
    return worst


def hostseam_lib():
    """g++ build of the engine's device physics (tests/hostseam): CPU-side test seam only."""
    d = os.path.join(HERE, "hostseam")
    so = os.path.join(d, "_build", "libemc_hostseam.so")
    srcs = [os.path.join(d, "hostseam.cpp"),
            os.path.join(os.path.dirname(HERE), "erpl_monte_carlo_sim_b200", "csrc", "emc_physics.cuh"),
            os.path.join(os.path.dirname(HERE), "erpl_monte_carlo_sim_b200", "csrc", "emc_model_build.h"),
            os.path.join(os.path.dirname(HERE), "erpl_monte_carlo_sim_b200", "csrc", "emc_strict.cuh"),
            os.path.join(os.path.dirname(HERE), "include", "emc.h")]
    if not os.path.isfile(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-std=c++17", "-ffp-contract=off",
                               "-o", so, srcs[0], "-lm"])
    return C.CDLL(so)


_dp = C.POINTER(C.c_double)


def hostseam_derivative(md, scalars, wind, t, state, chute):
    HS = hostseam_lib()
    m, keep = _abi.pack_model(md)
    scalars = np.ascontiguousarray(scalars, np.float64)
    n = scalars.shape[1]
    wind = np.ascontiguousarray(wind, np.float64) if wind is not None and np.size(wind) else None
    ins = _abi.inputs_struct(scalars, wind)
    ch = np.ascontiguousarray(chute, np.int32).copy()
    sd = np.empty((n, 14))
    t = np.ascontiguousarray(t, np.float64); state = np.ascontiguousarray(state, np.float64)
    rc = HS.hs_derivative(C.byref(m), C.byref(ins), C.c_int64(n), t.ctypes.data_as(_dp), state.ctypes.data_as(_dp),
                          ch.ctypes.data_as(C.POINTER(C.c_int32)), sd.ctypes.data_as(_dp))
    assert rc == 0
    return sd, ch


def hostseam_batch(md, scalars, wind, nan_ff=True):
    HS = hostseam_lib()
    m, keep = _abi.pack_model(md)
    scalars = np.ascontiguousarray(scalars, np.float64)
    n = scalars.shape[1]
    wind = np.ascontiguousarray(wind, np.float64) if wind is not None and np.size(wind) else None
    ins = _abi.inputs_struct(scalars, wind, wind_shared=(wind is not None and wind.ndim == 2))
    outs, out, iout = _abi.outputs_alloc(n)
    rc = HS.hs_batch(C.byref(m), C.byref(ins), C.c_int64(n), C.byref(outs), C.c_int(1 if nan_ff else 0), None,
                     C.c_int64(0), None)
    assert rc == 0
    return out, iout


def hostseam_batch_strict(md, scalars, wind, out_with_rail=None, nan_ff=True, tape=None):
    """Flights finished by the strict continuation (csrc/emc_strict.cuh) compiled by g++.  out_with_rail: an output block
    whose rail_* fields are filled (strict from the rail exit); None: fast path until the blow-up trigger, then strict.
    Returns (out, iout, strict_steps)."""
    HS = hostseam_lib()
    m, keep = _abi.pack_model(md)
    scalars = np.ascontiguousarray(scalars, np.float64)
    n = scalars.shape[1]
    wind = np.ascontiguousarray(wind, np.float64) if wind is not None and np.size(wind) else None
    ins = _abi.inputs_struct(scalars, wind, wind_shared=(wind is not None and wind.ndim == 2))
    outs, out, iout = _abi.outputs_alloc(n)
    if out_with_rail is not None:
        out[:] = out_with_rail
        iout[:] = 0
    steps = np.zeros(n, np.int64)
    rc = HS.hs_batch_strict(C.byref(m), C.byref(ins), C.c_int64(n), C.byref(outs), C.c_int(1 if nan_ff else 0),
                            C.c_int(-1 if out_with_rail is not None else 0), steps.ctypes.data_as(C.POINTER(C.c_int64)),
                            tape.ctypes.data_as(_dp) if tape is not None else None, C.c_int64(tape.shape[0] if tape is not None else 0))
    assert rc == 0
    return out, iout, steps


def single_case(z, name):
    md = _abi.model_from_npz(z, name + "__")
    wind = z[name + "__wind"][0] if (name + "__wind") in z.files else None
    return md, z[name + "__scalars"], wind, z[name + "__out"], z[name + "__iout"]


MC_SETS = ["mc_planar_liquid", "mc_planar_solid", "mc_liquid_default", "mc_solid_csv", "mc_readme_literal", "mc_solid_csv_blowup"]
DERIV_SETS = ["derivative_liquid_wind100", "derivative_solid_csv", "derivative_liquid_nowind"]


def synth(z, n, seed):
    """n seeded synthetic samples around a golden set: resample columns, jitter masses/thrust/attitude/wind."""
    rng = np.random.RandomState(seed)
    pick = rng.randint(0, z["scalars"].shape[1], n)
    sc = z["scalars"][:, pick].copy()
    wind = z["wind"][pick].copy()
    IN = _abi.IN
    k = rng.normal(1.0, 0.02, n)
    sc[IN["dry_mass"]] *= k; sc[IN["prop_mass"]] *= k
    sc[IN["burn_time"]] = sc[IN["prop_mass"]] / sc[IN["mdot"]]
    q = sc[IN["q0"]:IN["q3"] + 1] + rng.normal(0, 2e-3, (4, n))
    sc[IN["q0"]:IN["q3"] + 1] = q / np.linalg.norm(q, axis=0)
    sc[IN["vx"]:IN["vz"] + 1] += rng.normal(0, 0.1, (3, n))
    wind *= rng.uniform(0.5, 1.5, (n, 1, 1))
    return np.ascontiguousarray(sc), np.ascontiguousarray(wind)




# golden series keys (oracle/make_golden.py SERIES_KEYS) -> rows of the engine's series block
SERIES_MAP = {"mass": "mass", "center_of_mass": "center_of_mass", "thrust": "thrust", "drag": "drag", "cd": "cd", "cl": "cl",
              "cm": "cm", "cp_location_dynamic": "cp_location_dynamic", "stability_margin": "stability_margin",
              "angle_of_attack": "angle_of_attack", "sideslip_angle": "sideslip_angle", "speed": "speed"}


def series_reference(z, name):
    """(row indices, tape rows, {engine series field: reference values}) of one golden single flight."""
    idx = z[name + "__series__idx"]
    rows = z[name + "__series__tape"]
    ref = {v: z[name + "__series__" + k] for k, v in SERIES_MAP.items()}
    moi = z[name + "__series__moments_of_inertia"]; eul = z[name + "__series__euler_angles"]
    ref.update({"Ixx": moi[0], "Iyy": moi[1], "Izz": moi[2], "euler_roll": eul[0], "euler_pitch": eul[1], "euler_yaw": eul[2]})
    return idx, rows, ref


def assert_series_close(series, ref, what, rtol=RTOL):
    """Every derived series within rtol of the reference; angles get a 1e-9 rad floor, forces/coefficients are
    scaled by the series' own magnitude (they pass through zero); the blown-up tail (> 1e15) is skipped."""
    for field, rv in ref.items():
        got = series[_abi.SER[field]]
        ok = np.isfinite(rv) & (np.abs(rv) < 1e15)
        if not ok.any():
            continue
        scale = np.maximum(np.abs(rv[ok]), max(1e-9, 1e-6 * float(np.max(np.abs(rv[ok])))))
        if field in ("angle_of_attack", "sideslip_angle", "euler_roll", "euler_pitch", "euler_yaw"):
            # +pi and -pi are the same attitude
            d = np.abs(np.angle(np.exp(1j * (got[ok] - rv[ok]))))
            scale = np.maximum(np.abs(rv[ok]), 1e-3)
        else:
            d = np.abs(got[ok] - rv[ok])
        worst = int(np.argmax(d / scale))
        assert (d / scale)[worst] <= rtol, f"{what}: series {field} row {worst}: {got[ok][worst]!r} vs {rv[ok][worst]!r}"


def hostseam_series(md, scalars, wind, tape_rows):
    HS = hostseam_lib()
    m, keep = _abi.pack_model(md)
    scalars = np.ascontiguousarray(scalars, np.float64)
    wind = np.ascontiguousarray(wind, np.float64) if wind is not None and np.size(wind) else None
    ins = _abi.inputs_struct(scalars, wind, wind_shared=True)
    tape_rows = np.ascontiguousarray(tape_rows, np.float64)
    n = tape_rows.shape[0]
    out = np.empty((_abi.SERIES_COUNT, n))
    rc = HS.hs_series(C.byref(m), C.byref(ins), tape_rows.ctypes.data_as(_dp), C.c_int64(n), out.ctypes.data_as(_dp))
    assert rc == 0
    return out


def hostseam_component(md, comp, cols):
    HS = hostseam_lib()
    m, keep = _abi.pack_model(md)
    n_out = {0: 5, 1: 5, 2: 7, 3: 1}[comp]
    blk = np.ascontiguousarray(np.stack([np.asarray(c, np.float64) for c in np.broadcast_arrays(*cols)], axis=0))
    n = blk.shape[1]
    out = np.empty((n_out, n))
    rc = HS.hs_component(C.byref(m), C.c_int(comp), C.c_int64(n), blk.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
    assert rc == 0
    return out


def check_components(evaluate, z):
    """Component known-answer tests against the reference (tests/golden/components.npz); `evaluate(kind, comp, cols)`
    runs the engine's component functions for the 'liquid' or 'solid' model."""
    o = evaluate("liquid", 0, (z["atm_z"],))
    np.testing.assert_allclose(o[:3].T, z["atm_out"], rtol=4e-15)
    np.testing.assert_allclose(o[4], z["gravity"], rtol=2e-15)
    o = evaluate("liquid", 1, (z["mp_pf"], 113.4 * z["mp_mult"], 63.5 * z["mp_mult"]))
    np.testing.assert_allclose(o[:4].T, z["mp_out"], rtol=1e-15)
    o = evaluate("liquid", 2, (z["aero_mach"], z["aero_alpha"], z["aero_beta"], z["aero_cg"], z["aero_power_on"].astype(float), 1.0))
    np.testing.assert_allclose(o[[0, 1, 2, 3, 4, 5]].T, z["aero_out"], rtol=2e-14, atol=1e-16)
    ls, ss = z["liquid_scalars"], z["solid_scalars"]
    o = evaluate("liquid", 3, (z["thr_t"], z["thr_p"], ls[0], ls[1], ls[3]))
    np.testing.assert_allclose(o[0], z["thr_liquid"], rtol=1e-15)
    o = evaluate("solid", 3, (z["thr_t"], z["thr_p"], ss[0], ss[1], ss[3]))
    np.testing.assert_allclose(o[0], z["thr_solid"], rtol=1e-15)
