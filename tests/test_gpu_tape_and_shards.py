"""GPU (-m gpu): round-2 boundary features — the downsampled batch tape (reference monte_carlo.py:296-302 'trajectory'),
the plot methods (monte_carlo.py:562-707), several contexts on one device, sample sharding of run_monte_carlo over the
ranks of a torch.distributed job (the fan-out of monte_carlo.py:63-83) with NCCL-reduced statistics."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import util
from erpl_monte_carlo_sim_b200 import (LiquidMotor, MonteCarloAnalyzer, Rocket, SolidMotor, StandardAtmosphere, WindModel,
                                       _abi, _lib, stats)
from test_host_sampling import CSV_ALT, CSV_WIND

pytestmark = pytest.mark.gpu
VERTICAL = [0.0, -np.pi / 2 + 0.02, 0.0]
IC = {"position": [0.0, 0.0, 10.0], "velocity": [0, 0, 0.0], "attitude": VERTICAL, "angular_velocity": [0.0, 0.0, 0.0]}


def _csv_analyzer():
    mc = MonteCarloAnalyzer(Rocket(), SolidMotor(), StandardAtmosphere(), WindModel())
    mc.base_altitude_profile, mc.base_wind_profile = CSV_ALT, CSV_WIND
    return mc


def test_batch_tape_matches_the_full_tape(engine):
    """Rows of the downsampled batch tape are the stored states 0, k, 2k, ... and the last one, bit for bit what the
    one-flight tape (every stored state) holds; arming the tape does not change the summaries."""
    z = util.golden("mc_solid_csv")
    engine.set_model(_abi.model_from_npz(z))
    sc, wind = z["scalars"], z["wind"]
    plain_out, plain_iout = engine.run_batch(sc, wind)
    picks = np.array([0, 3, 5, 17, 63], np.int64)
    for stride in (7, 20):
        cap = 60000 // stride + 4
        engine.tape_request(picks, stride, cap)
        out, iout = engine.run_batch(sc, wind)
        assert engine.counters()["tape_rows"] > 0
        rows, cnt = engine.tape_fetch()
        np.testing.assert_array_equal(iout, plain_iout)
        np.testing.assert_array_equal(out, plain_out)                       # NaN-aware equality of the summaries
        for k, i in enumerate(picks):
            o1, io1, full = engine.run_tape(sc[:, i:i + 1], wind[i])
            t_rail = o1[_abi.OUT["rail_exit_time"], 0]
            n_steps, first_nan = int(iout[_abi.IOUT["n_steps"], i]), int(iout[_abi.IOUT["first_nan_step"], i])
            m = int(cnt[k])
            assert 2 <= m <= cap
            strided = np.arange(0, m - 1) * stride
            got = rows[k, :m]
            want = np.column_stack([full[strided, 0] - t_rail, full[strided, 1:4]])
            np.testing.assert_array_equal(got[:m - 1], want)
            if first_nan < 0:                                                # flown to its end: the last row is the final state
                assert m == n_steps // stride + 1 + (1 if n_steps % stride else 0)
                np.testing.assert_array_equal(got[m - 1], np.concatenate([[full[n_steps, 0] - t_rail], full[n_steps, 1:4]]))
    # a request is consumed by one run
    engine.run_batch(sc, wind)
    assert engine.counters()["tape_rows"] == 0


def test_one_flight_paths_leave_the_resident_outputs_alone(engine):
    """simulate_flight / full_result (tape + series of ONE flight) run between a batch and its statistics: the batch's
    outputs in HBM must survive (round-1 advisor finding)."""
    z = util.golden("mc_liquid_default")
    engine.set_model(_abi.model_from_npz(z))
    out, iout = engine.run_batch(z["scalars"], z["wind"])
    before = stats.device_statistics(engine, out.shape[1])
    _, _, tape = engine.run_tape(z["scalars"][:, 2:3], z["wind"][2])
    engine.extract_series(z["scalars"][:, 2:3], z["wind"][2], tape)
    engine.derivative_debug(z["scalars"][:, :4], z["wind"][:4], np.zeros(4), np.tile(tape[5, 1:], (4, 1)), np.zeros(4, np.int32))
    after = stats.device_statistics(engine, out.shape[1])
    assert json.dumps(before, sort_keys=True, default=float) == json.dumps(after, sort_keys=True, default=float)


def test_two_contexts_on_one_device_keep_their_own_model(engine):
    """The run constants live in per-device __constant__ memory: a second context with another model must not change
    what the first one flies (round-1 finding: emc_engine.cu c_model / c_tables)."""
    za, zb = util.golden("mc_liquid_default"), util.golden("mc_solid_csv")
    e1, e2 = _lib.Engine(0), _lib.Engine(0)
    try:
        e1.set_model(_abi.model_from_npz(za))
        e2.set_model(_abi.model_from_npz(zb))                 # uploads ITS constants to the shared bank
        for _ in range(2):                                    # interleaved launches
            oa, ia = e1.run_batch(za["scalars"], za["wind"])
            ob, ib = e2.run_batch(zb["scalars"], zb["wind"])
            np.testing.assert_array_equal(ia, za["iout"]); np.testing.assert_array_equal(ib, zb["iout"])
            util.assert_summary_close(oa, za["out"], what="context 1 (liquid, 100-knot wind)")
            util.assert_summary_close(ob, zb["out"], what="context 2 (solid, CSV wind)")
        c1 = e1.component(0, np.array([5000.0]))              # component evaluation re-asserts residency too
        assert abs(c1[1, 0] - 54019.90357580142) < 1e-9 * 54019.9
    finally:
        e1.close(); e2.close()


def test_results_carry_trajectories_and_the_plot_methods_run(tmp_path, monkeypatch):
    """example.py:57-66: run_monte_carlo, then plot_results / plot_trajectory_cloud(_3d).  Every result serves
    'trajectory' {time, altitude, position(n,3)} (monte_carlo.py:296-302) from the kernel's downsampled tape; samples
    the run did not tape are taped by one extra batch; all equal the full one-flight tape at the stored rows."""
    import mpl_stub
    mc = _csv_analyzer()
    mc.trajectory_samples, mc.trajectory_stride = 24, 10
    for k in ("initial_velocity", "initial_attitude", "initial_angular_velocity"):       # planar: flights reach landing
        mc.uncertainty_params[k] = [0.0, 0.0, 0.0]
    mc.uncertainty_params["initial_attitude"] = [0.0, 0.005, 0.0]
    mc.uncertainty_params["wind_speed_range"] = [0.0, 0.0]
    mc.wind_model.turbulence_intensity = 0.0
    mc.base_wind_profile = CSV_WIND * np.array([1.0, 0.0, 0.0])
    an = mc.run_monte_carlo(IC, n_samples=64)
    run = mc.last_run
    assert an["n_samples"] >= 32 and run.tape_ids.size == 24
    res = an["results"]
    r0 = res[0]
    assert "trajectory" in r0 and set(r0["trajectory"]) == {"time", "altitude", "position"}
    tr = r0["trajectory"]
    assert tr["position"].shape == (tr["time"].size, 3) and np.array_equal(tr["altitude"], tr["position"][:, 2])
    assert tr["time"][0] == 0.0 and abs(tr["time"][-1] - r0["flight_time"]) <= 1e-12 * r0["flight_time"]
    full = run.full_result(int(r0["simulation_id"]))
    k = np.arange(tr["time"].size - 1) * mc.trajectory_stride
    np.testing.assert_allclose(tr["position"][:-1], full["trajectory"]["position"][k], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(tr["time"][:-1], full["trajectory"]["time"][k], rtol=0, atol=1e-12)
    late = res[len(res) - 1]                                                # beyond the taped prefix: taped on demand
    assert int(late["simulation_id"]) >= 24 and late["trajectory"]["altitude"].max() > 1000.0
    assert abs(late["trajectory"]["altitude"].max() - late["apogee_altitude"]) < 5.0
    monkeypatch.chdir(tmp_path)
    calls = mpl_stub.install()
    try:
        out_dir = mc.plot_results(an)
        assert os.path.isfile(os.path.join(out_dir, "monte_carlo_distributions.png"))
        assert os.path.isfile(os.path.join(out_dir, "monte_carlo_report.json"))
        assert sum(1 for c in calls if c[1] == "hist") == 3 and sum(1 for c in calls if c[1] == "scatter") == 1
        del calls[:]
        mc.plot_trajectory_cloud(an, max_trajectories=40)
        n_plot = sum(1 for c in calls if c[1] == "plot")
        assert n_plot == 2 * min(40, an["n_samples"])
        del calls[:]
        mc.plot_trajectory_cloud_3d(an, save_plots=False, max_trajectories=8)
        assert sum(1 for c in calls if c[0] == "axes3d" and c[1] == "plot") == min(8, an["n_samples"])
    finally:
        mpl_stub.remove()
    with pytest.raises(ImportError, match="matplotlib"):
        mc.plot_results(an, save_plots=False)


def test_device_rng_run_serves_trajectories():
    mc = _csv_analyzer()
    mc.rng = "numpy-device"
    mc.trajectory_samples = 4
    an = mc.run_monte_carlo(IC, n_samples=300)
    run = mc.last_run
    host = _csv_analyzer(); host.rng = "numpy"; host.trajectory_samples = 0
    ref = host.run_batch(IC, host.draw_parameters(300))
    np.testing.assert_array_equal(run.iout[_abi.IOUT["rail_steps"]], ref.iout[_abi.IOUT["rail_steps"]])
    same = (run.iout == ref.iout).all(axis=0)
    assert same.mean() > 0.97                                  # device-regenerated normals differ in the last place for a few
    res = an["outliers"]
    a, b = res[0], res[len(res) - 1]
    assert a["trajectory"]["time"].size >= 2 and b["trajectory"]["time"].size >= 2
    assert run.scalars.shape == (_abi.IN_COUNT, 300) and run.disp.n == 300


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_device_generated_batch_keeps_its_outputs_in_hbm_until_read():
    """A device-generated campaign returns its statistics without moving the per-sample outputs (emc_fetch_outputs on
    access).  A second campaign on the same engine must not lose the first one's results: the engine downloads a pending
    batch before it is overwritten."""
    mc = _csv_analyzer()
    mc.rng = "numpy-device"; mc.trajectory_samples = 0
    a1 = mc.run_monte_carlo(IC, n_samples=3000)
    run1 = mc.last_run
    assert run1._out is None, "the outputs should still be in HBM only"
    assert len(a1["results"]) == a1["n_samples"] and run1._out is None          # the length comes from the device statistics
    mc2 = _csv_analyzer()
    mc2.rng = "numpy-device"; mc2.trajectory_samples = 0
    a2 = mc2.run_monte_carlo(IC, n_samples=2000)
    assert run1._out is not None, "the second run must have spilled the first run's outputs to the host"
    # the same seeds flown with host draws give the same per-sample summaries (inputs differ by <= 2 ulp in 0.1 % of the normals)
    mc3 = _csv_analyzer()
    ref = mc3.run_batch(IC, mc3.draw_parameters(3000))
    same = np.all(run1.iout == ref.iout, axis=0)
    assert same.mean() > 0.995
    ok = np.isfinite(ref.out[_abi.OUT["apogee_altitude"]]) & same
    np.testing.assert_allclose(run1.out[_abi.OUT["apogee_altitude"]][ok], ref.out[_abi.OUT["apogee_altitude"]][ok], rtol=1e-6)
    r0 = a1["results"][0]
    assert r0["apogee_altitude"] == run1.out[_abi.OUT["apogee_altitude"], a1["results"].sample_indices[0]]
    assert len(a2["results"]) == a2["n_samples"]
    assert [len(a1["results"]), len(a1["outliers"])] == [a1["n_samples"], a1["n_outliers"]]


def _raw_summary(engine, n, pcts):
    import ctypes as C
    pct = (C.c_double * len(pcts))(*[float(p) for p in pcts])
    res = np.empty(32 + 6 * len(pcts))
    engine._check(engine._lib.emc_stats_summary(engine._ctx, None, 0, n, pct, len(pcts), res.ctypes.data_as(C.POINTER(C.c_double))), "emc_stats_summary")
    return res


@pytest.mark.parametrize("n_dev", [1, 2])
def test_device_group_through_the_c_abi(engine, n_dev):
    """emc_group (torch-free multi-device path of the C ABI): the shards flown by the group are bit-identical to one
    engine flying everything, and the NCCL-reduced statistics equal the one-device statistics of the whole batch."""
    import torch
    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    z = util.golden("mc_solid_csv")
    md = _abi.model_from_npz(z)
    sc, wind = util.synth(z, 3001, seed=5)                     # odd count: uneven shards
    sc, wind = np.ascontiguousarray(sc), np.ascontiguousarray(wind)
    engine.set_model(md)
    ref_out, ref_iout = engine.run_batch(sc, wind)
    pcts = (5, 25, 50, 75, 95)
    ref_stats = _raw_summary(engine, sc.shape[1], pcts)
    grp = _lib.EngineGroup(list(range(n_dev)))
    try:
        grp.set_model(md)
        out, iout = grp.run_batch(sc, wind)
        np.testing.assert_array_equal(iout, ref_iout)
        np.testing.assert_array_equal(out, ref_out)            # same kernels, same samples: bit for bit (NaN == NaN by position)
        sh = grp.shards()
        assert sh[0][0] == 0 and sum(c for _, c in sh) == sc.shape[1] and all(sh[i][0] + sh[i][1] == sh[i + 1][0] for i in range(n_dev - 1))
        st = grp.stats_summary(pcts)
        # counts, min / max and the order statistics are exact; the sums differ by the order of summation
        np.testing.assert_array_equal(st[:9], ref_stats[:9])
        np.testing.assert_array_equal(st[14:20], ref_stats[14:20])
        np.testing.assert_array_equal(st[32:], ref_stats[32:])
        np.testing.assert_allclose(st[9:14], ref_stats[9:14], rtol=1e-12)
        np.testing.assert_allclose(st[20:31], ref_stats[20:31], rtol=1e-9)
        c = grp.counters()
        assert c["refills"] == sc.shape[1] and c["kernel_launches"] == 4 * n_dev
    finally:
        grp.close()


def test_sharded_run_monte_carlo_equals_single_rank(tmp_path):
    """torchrun x2: rank r flies seeds [r n/2, (r+1) n/2) and every rank reports the statistics of the whole job
    (NCCL all-reduce between the passes); equal to the one-process run on the same seeds, statistics included."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = 3001
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dist_shard_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), worker, str(tmp_path), str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    mc = _csv_analyzer(); mc.rng = "numpy"; mc.host_rng_max = 1 << 30
    an = mc.run_monte_carlo(IC, n_samples=n)
    one = mc.last_run
    parts = [np.load(os.path.join(str(tmp_path), f"rank{k}.npz"), allow_pickle=True) for k in range(2)]
    assert int(parts[0]["first_id"]) == 0 and int(parts[1]["first_id"]) == n // 2
    np.testing.assert_array_equal(np.concatenate([p["iout"] for p in parts], axis=1), one.iout)
    np.testing.assert_array_equal(np.concatenate([p["out"] for p in parts], axis=1), one.out)
    for p in parts:
        got = json.loads(str(p["analysis"]))
        assert got["n_samples"] == an["n_samples"] and got["n_outliers"] == an["n_outliers"]
        for key in ("apogee_altitude", "range", "flight_time"):
            for f in ("mean", "std", "min", "max"):
                assert abs(got[key][f] - an[key][f]) <= 1e-12 * abs(an[key][f])
            np.testing.assert_allclose(got[key]["percentiles"], an[key]["percentiles"], rtol=1e-15)
        assert got["parameter_ranges_observed"] == json.loads(json.dumps(an["parameter_ranges_observed"]))
