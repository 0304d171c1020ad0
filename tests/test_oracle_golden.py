"""CPU: pins the C oracle (oracle/emc_oracle.c) to golden vectors produced by the unmodified Python
reference (oracle/make_golden.py).  Integer results (step counts, termination codes, apogee indices)
must be identical; floating-point summaries within 1e-6 relative (observed <= 1e-7, and that only on
the reference's super-exponentially diverging flights)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import util
from erpl_monte_carlo_sim_b200 import _abi


@pytest.mark.parametrize("name", util.DERIV_SETS)
def test_oracle_derivative(name):
    z = util.golden(name)
    md = _abi.model_from_npz(z)
    sd, ch = O.derivative(md, z["scalars"], z["wind"] if z["wind"].size else None, z["t"], z["state"], z["chute_in"])
    ref = z["state_dot"]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.nanmax(np.abs(sd - ref) / np.maximum(scale, 1e-300)) < 1e-13
    assert np.array_equal(np.isnan(sd), np.isnan(ref))
    assert np.array_equal(ch, z["chute_out"])


def test_oracle_components():
    z = util.golden("components")
    L = O.lib()
    m, keep = _abi.pack_model(_abi.model_from_npz(z, "liquid_"))
    ms, keeps = _abi.pack_model(_abi.model_from_npz(z, "solid_"))
    T, p, r = C.c_double(), C.c_double(), C.c_double()
    got = []
    for zz in z["atm_z"]:
        L.orc_atmosphere(C.byref(m), zz, C.byref(T), C.byref(p), C.byref(r))
        got.append([T.value, p.value, r.value])
    np.testing.assert_allclose(np.array(got), z["atm_out"], rtol=2e-15)
    g = np.array([L.orc_gravity(C.byref(m), zz) for zz in z["atm_z"]])
    np.testing.assert_array_equal(g, z["gravity"])
    mp = (C.c_double * 4)()
    got = []
    for f, k in zip(z["mp_pf"], z["mp_mult"]):
        L.orc_mass_properties(C.byref(m), 113.4 * k, 63.5 * k, f, mp)
        got.append(list(mp))
    np.testing.assert_array_equal(np.array(got), z["mp_out"])
    c = (C.c_double * 6)()
    got = []
    for M, a, b, cg, po in zip(z["aero_mach"], z["aero_alpha"], z["aero_beta"], z["aero_cg"], z["aero_power_on"]):
        L.orc_aero_coefficients(C.byref(m), M, a, b, cg, int(po), 1.0, c)
        got.append(list(c))
    np.testing.assert_allclose(np.array(got), z["aero_out"], rtol=1e-15, atol=0)
    ls, ss = z["liquid_scalars"], z["solid_scalars"]
    tl = [L.orc_thrust(C.byref(m), ls[0], ls[1], ls[3], t, pp) for t, pp in zip(z["thr_t"], z["thr_p"])]
    ts = [L.orc_thrust(C.byref(ms), ss[0], ss[1], ss[3], t, pp) for t, pp in zip(z["thr_t"], z["thr_p"])]
    np.testing.assert_array_equal(np.array(tl), z["thr_liquid"])
    np.testing.assert_array_equal(np.array(ts), z["thr_solid"])
    dp = C.POINTER(C.c_double)
    xp = np.array(_abi.model_from_npz(z, "liquid_")["cd_mach"]); fp = np.array(_abi.model_from_npz(z, "liquid_")["cd0"])
    it = np.array([L.orc_interp(x, xp.ctypes.data_as(dp), fp.ctypes.data_as(dp), xp.size) for x in z["interp_x"]])
    np.testing.assert_array_equal(it, z["interp_out"])
    q = (C.c_double * 4)()
    qs = []
    for e in z["euler"]:
        L.orc_euler_to_quaternion(*e, q)
        qs.append(list(q))
    np.testing.assert_array_equal(np.array(qs), z["quat"])
    e3 = (C.c_double * 3)(); R9 = (C.c_double * 9)()
    for qr, eb, rot in zip(z["quat_raw"], z["euler_back"], z["rot"]):
        qq = (C.c_double * 4)(*qr)
        L.orc_quaternion_to_euler(qq, e3)
        np.testing.assert_allclose(np.array(list(e3)), eb, rtol=1e-15, atol=1e-16)
        L.orc_rotation_matrix(qq, R9)
        np.testing.assert_allclose(np.array(list(R9)).reshape(3, 3), rot, rtol=0, atol=2e-15)


def test_oracle_single_flights():
    z = util.golden("flights_single")
    for name in z["names"]:
        md, sc, wind, ref, iref = util.single_case(z, str(name))
        out, iout = O.batch(md, sc, wind)
        np.testing.assert_array_equal(iout, iref, err_msg=str(name))
        util.assert_summary_close(out, ref, what=str(name))


@pytest.mark.parametrize("name", util.MC_SETS)
def test_oracle_mc_sets(name):
    z = util.golden(name)
    md = _abi.model_from_npz(z)
    out, iout = O.batch(md, z["scalars"], z["wind"])
    np.testing.assert_array_equal(iout, z["iout"])
    util.assert_summary_close(out, z["out"], what=name)


def test_oracle_tape_matches_reference_states():
    z = util.golden("flights_single")
    name = "c1b_example_liquid_csv"
    md, sc, wind, ref, iref = util.single_case(z, name)
    out, iout, tape = O.tape(md, sc, wind)
    assert tape.shape[0] == iref[0, 0] + 1
    idx = z[name + "__series__idx"]
    ref_rows = z[name + "__series__tape"]
    ok = np.abs(ref_rows) < 1e15                      # the diverged tail is compared by the summary test
    np.testing.assert_allclose(tape[idx][ok], ref_rows[ok], rtol=1e-6, atol=1e-9)


def test_oracle_threads_agree():
    z = util.golden("mc_solid_csv")
    md = _abi.model_from_npz(z)
    a = O.batch(md, z["scalars"][:, :16], z["wind"][:16], n_threads=1)
    b = O.batch(md, z["scalars"][:, :16], z["wind"][:16], n_threads=4)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_oracle_series_match_reference():
    """_extract_results (simulator.py:496-552) restated in the oracle vs the reference's own series, evaluated on the
    reference's stored states (rows of the golden tape), incl. the shifted-time thrust quirk (:543)."""
    z = util.golden("flights_single")
    for name in z["names"]:
        name = str(name)
        md, sc, wind, ref, iref = util.single_case(z, name)
        idx, rows, sref = util.series_reference(z, name)
        rows = rows.copy()
        # row 0 of the selection is state 0 (idx[0] == 0): the oracle takes t_rail from the first row
        assert idx[0] == 0
        got = O.series(md, sc, wind, rows)
        util.assert_series_close(got, sref, name, rtol=1e-9)
