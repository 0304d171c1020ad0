"""Device-side dispersion draws (csrc/emc_philox.cuh).
CPU: Philox4x32-10 known-answer vectors (Random123 kat_vectors) for the host mirror.
GPU: device draws == host mirror (uniform bits exact); the device perturbation fed with the reference's own NumPy
draws reproduces the host-seeded inputs; Philox-drawn inputs have the reference's distribution (KS tests) and its
stream structure; runs are reproducible and independent of sharding."""
import numpy as np
import pytest

import util
from erpl_monte_carlo_sim_b200 import (LiquidMotor, MonteCarloAnalyzer, Rocket, SolidMotor, StandardAtmosphere, WindModel, _abi,
                                       philox)
from test_host_sampling import CSV_ALT, CSV_WIND

VERTICAL = [0.0, -np.pi / 2 + 0.02, 0.0]


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        r = philox.philox4x32_10([c[0]], [c[1]], [c[2]], [c[3]], k[0], k[1])
        assert tuple(int(x[0]) for x in r) == want


def _mc(solid, csv):
    mc = MonteCarloAnalyzer(Rocket(), SolidMotor() if solid else LiquidMotor(), StandardAtmosphere(), WindModel())
    if csv:
        mc.base_altitude_profile, mc.base_wind_profile = CSV_ALT, CSV_WIND
    return mc


@pytest.mark.gpu
def test_device_draws_equal_host_mirror(engine):
    g, u = engine.philox_draws(seed=0x1234567890abcdef, first_index=(1 << 33) + 5, n=4096, n_gauss=31)
    idx = np.arange((1 << 33) + 5, (1 << 33) + 5 + 4096, dtype=np.uint64)
    np.testing.assert_array_equal(u, philox.uniforms(0x1234567890abcdef, idx))        # same bits
    np.testing.assert_allclose(g, philox.normals(0x1234567890abcdef, idx, 31), rtol=0, atol=2e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("solid,csv,name", [(False, False, "mc_liquid_default"), (True, True, "mc_solid_csv")])
def test_device_perturbation_reproduces_host_seeded_inputs(engine, solid, csv, name):
    """Feed the reference's own NumPy draws through the device generator: inputs equal the golden inputs."""
    z = util.golden(name)
    mc = _mc(solid, csv)
    n = z["scalars"].shape[1]
    n_knots = z["wind"].shape[1]
    rs = np.random.RandomState(0)
    G = max(15, 3 * n_knots)
    gauss = np.empty((n, G)); unif = np.empty((n, 2))
    for i in range(n):                                         # the stream of seed i: 14 normals, 2 uniforms, 1 normal ...
        rs.seed(i); a = rs.standard_normal(14); unif[i] = rs.random_sample(2); b = rs.standard_normal()
        rs.seed(i); gauss[i] = rs.standard_normal(G)           # ... and the same stream again for motor and wind (F11)
        assert np.array_equal(gauss[i, :14], a)
        if G == 15:
            gauss[i, 14] = b
    b = z["base_ic"]
    ic = {"position": b[0], "velocity": b[1], "attitude": b[2], "angular_velocity": b[3]}
    engine.set_model(_abi.model_from_npz(z))
    engine.generate_inputs(mc.dispersion_struct(ic), 0, 0, n, gauss=gauss, unif=unif)
    sc, wind = engine.staged_inputs(n)
    np.testing.assert_allclose(sc, z["scalars"], rtol=3e-15, atol=1e-16)
    np.testing.assert_allclose(wind, z["wind"], rtol=1e-13, atol=1e-13)
    out, iout = engine.run_batch_staged(n)
    np.testing.assert_array_equal(iout, z["iout"])
    util.assert_summary_close(out, z["out"], what="staged " + name, sens=util.oracle_sensitivity(_abi.model_from_npz(z), z["scalars"], z["wind"]))


@pytest.mark.gpu
def test_philox_inputs_have_the_reference_distribution(engine):
    from scipy import stats as sst
    mc = _mc(False, False)
    mc.philox_seed = 20240601
    n = 20000
    ic = {"attitude": VERTICAL}
    engine.set_model(__import__("erpl_monte_carlo_sim_b200").marshal.model_dict(mc.rocket, mc.motor, mc.atmosphere, mc._model_simulator(),
                                                                                 mc._altitude_grid()))
    engine.generate_inputs(mc.dispersion_struct(ic), mc.philox_seed, 0, n)
    sc, wind = engine.staged_inputs(n)
    ref_sc, ref_wind, _ = mc.build_inputs(ic, mc.draw_parameters(n))          # the reference's MT19937 draws
    IN = _abi.IN
    for f in ("vx", "vy", "vz", "wx", "dry_mass", "thrust_a", "mdot", "burn_time", "nozzle_area"):
        assert sst.ks_2samp(sc[IN[f]], ref_sc[IN[f]]).pvalue > 1e-3, f
    for k in (0, 1, 50, 99):
        for c in range(3):
            assert sst.ks_2samp(wind[:, k, c], ref_wind[:, k, c]).pvalue > 1e-3, (k, c)
    # F11 stream structure: the thrust multiplier and the surface gust share the sample's first Gaussian
    g0 = (sc[IN["thrust_a"]] / mc.motor.thrust_vacuum - 1.0) / mc.motor.thrust_uncertainty
    assert abs(np.corrcoef(g0, wind[:, 0, 0] - 0.0)[0, 1]) > 0.95
    assert abs(np.corrcoef((ref_sc[IN["thrust_a"]] / mc.motor.thrust_vacuum - 1.0), ref_wind[:, 0, 0])[0, 1]) > 0.95
    q = sc[IN["q0"]:IN["q3"] + 1]
    np.testing.assert_allclose(np.sum(q * q, axis=0), 1.0, atol=1e-14)
    p = mc.philox_parameters(64)
    np.testing.assert_allclose(sc[IN["dry_mass"], :64], mc.rocket.dry_mass * p.mass_multiplier, rtol=1e-13)


@pytest.mark.gpu
def test_philox_monte_carlo_end_to_end():
    mc = _mc(True, True)
    mc.rng, mc.philox_seed = "philox", 99
    ic = {"position": [0.0, 0.0, 10.0], "attitude": VERTICAL}
    a = mc.run_batch_philox(ic, 3000)
    b = mc.run_batch_philox(ic, 3000)
    assert np.array_equal(a.out, b.out, equal_nan=True) and np.array_equal(a.iout, b.iout)      # reproducible
    lo = mc.run_batch_philox(ic, 1000, first_index=0); hi = mc.run_batch_philox(ic, 2000, first_index=1000)
    assert np.array_equal(np.concatenate([lo.out, hi.out], 1), a.out, equal_nan=True)           # shard invariant
    mc.chunk_size = 700                                                                          # chunk invariant
    c = mc.run_batch_philox(ic, 3000)
    assert np.array_equal(c.out, a.out, equal_nan=True)
    mc.chunk_size = 1 << 16
    ref = mc.run_batch(ic, mc.draw_parameters(3000))           # host-seeded run: same distribution of outcomes
    steps_p, steps_h = a.iout[0].astype(float), ref.iout[0].astype(float)
    assert abs(np.median(steps_p[steps_p < 10000]) / np.median(steps_h[steps_h < 10000]) - 1) < 0.03
    an = mc.run_monte_carlo(ic, n_samples=2000)
    assert an["n_samples"] + an["n_outliers"] == 2000 and an["results"][0]["parameters"]["random_seed"] >= 0


@pytest.mark.gpu
def test_device_regenerates_numpy_streams(engine):
    """MT19937 + NumPy's legacy Gaussian on the device: uniforms bit-identical, normals within 4 ulp of NumPy's
    (CUDA log/sqrt vs libm), for small and large seeds and more than one block of 624 words."""
    seeds = [0, 1, 2, 41, 99999, 123456789]
    for first in seeds:
        g, u, dens = engine.numpy_draws(first, 3, 300)
        for i in range(3):
            rs = np.random.RandomState(first + i)
            a = rs.standard_normal(14); uu = rs.random_sample(2); dd = rs.standard_normal()
            ref = np.random.RandomState(first + i).standard_normal(300)
            np.testing.assert_array_equal(u[i], uu)
            assert np.max(np.abs(g[i] - ref) / np.spacing(np.abs(ref))) <= 4.0
            assert abs(dens[i] - dd) <= 4.0 * np.spacing(abs(dd))
            np.testing.assert_array_equal(ref[:14], a)


@pytest.mark.gpu
@pytest.mark.parametrize("solid,csv,name", [(False, False, "mc_liquid_default"), (True, True, "mc_solid_csv")])
def test_numpy_device_mode_matches_host_seeded_goldens(engine, solid, csv, name):
    """rng = "numpy-device": inputs regenerated on the GPU equal the reference's host-seeded inputs to rounding and the
    flights match the reference's goldens."""
    z = util.golden(name)
    mc = _mc(solid, csv)
    n = z["scalars"].shape[1]
    b = z["base_ic"]
    ic = {"position": b[0], "velocity": b[1], "attitude": b[2], "angular_velocity": b[3]}
    engine.set_model(_abi.model_from_npz(z))
    engine.generate_inputs_numpy(mc.dispersion_struct(ic), 0, n)
    sc, wind = engine.staged_inputs(n)
    np.testing.assert_allclose(sc, z["scalars"], rtol=5e-15, atol=1e-16)
    np.testing.assert_allclose(wind, z["wind"], rtol=1e-12, atol=1e-13)
    run = mc.run_batch_numpy_device(ic, n)
    np.testing.assert_array_equal(run.iout, z["iout"])
    util.assert_summary_close(run.out, z["out"], what="numpy-device " + name,
                              sens=util.oracle_sensitivity(_abi.model_from_npz(z), z["scalars"], z["wind"]))
    np.testing.assert_allclose(_param_matrix(run.disp), z["params"], rtol=5e-15, atol=1e-17)


def _param_matrix(d):
    return np.column_stack([d.pos, d.vel, d.att, d.omega, d.mass_multiplier, d.thrust_multiplier, d.wind_speed, d.wind_direction,
                            d.density_multiplier])


@pytest.mark.gpu
def test_device_drawn_c4_flights_match_the_oracle(engine):
    """BASELINE config C4 as the bench flies it (LiquidMotor, default dispersions, 100-knot stochastic wind per sample, drawn
    and perturbed on the device): the staged inputs are read back and the C oracle flies the same samples — integer outputs
    exact on every sample, summaries under the usual rule (tools/parity_sweep.py runs the same check on 30 000)."""
    import oracle_lib as O
    from erpl_monte_carlo_sim_b200 import marshal
    mc = MonteCarloAnalyzer(Rocket(), LiquidMotor(), StandardAtmosphere(), WindModel())
    ic = {"position": [0.0, 0.0, 0.0], "velocity": [0.0, 0.0, 0.0], "attitude": VERTICAL, "angular_velocity": [0.0, 0.0, 0.0]}
    md = marshal.model_dict(mc.rocket, mc.motor, mc.atmosphere, mc._model_simulator(), mc._altitude_grid())
    n = 1500
    engine.set_model(md)
    engine.generate_inputs(mc.dispersion_struct(ic), 7, 4_000_000, n)
    sc, wind = engine.staged_inputs(n, want_wind=True)
    assert wind.shape == (n, 100, 3)
    out, iout = engine.run_batch_staged(n)
    ref, iref = O.batch(md, sc, wind)
    np.testing.assert_array_equal(iout, iref)
    sens = util.oracle_sensitivity(md, sc, wind)
    # valid (non-outlier) flights: the usual rule.  Every 3-D flight of the reference ends in a super-exponential blow-up
    #  that amplifies one ulp without bound, and the three-input sensitivity estimate is only a lower bound
    # there (47 outliers of 30 000 C4 samples exceed 10x of it, by errors of 1e-6..7e-5): the outliers
    # are held to 1e-3, by category where non-finite or beyond 1e150.
    OUT = _abi.OUT
    bad = MonteCarloAnalyzer.outlier_mask(ref[OUT["apogee_altitude"]], ref[OUT["range"]], ref[OUT["flight_time"]])
    assert np.array_equal(bad, MonteCarloAnalyzer.outlier_mask(out[OUT["apogee_altitude"]], out[OUT["range"]], out[OUT["flight_time"]]))
    assert (~bad).sum() >= 10
    util.assert_summary_close(out[:, ~bad], ref[:, ~bad], what="device-drawn C4 flights", sens=sens[:, ~bad])
    err = util.summary_errors(out[:, bad], ref[:, bad])
    assert np.max(err) <= 1e-3          # a NaN / inf category mismatch is an infinite error
