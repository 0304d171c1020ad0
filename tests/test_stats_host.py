"""CPU: the statistics pass sequencing (erpl_monte_carlo_sim_b200/stats.py) — classification, two-pass moments,
exact np.percentile by radix select — against NumPy / the reference's _analyze_results golden, on one process and
on two `gloo` ranks that each hold half of the samples (the multi-GPU path with NCCL swapped for gloo)."""
import os
import socket

import numpy as np
import pytest

import util
from erpl_monte_carlo_sim_b200 import MonteCarloAnalyzer, Rocket, LiquidMotor, StandardAtmosphere, WindModel
from erpl_monte_carlo_sim_b200 import stats as S
from stats_numpy_backend import NumpyBackend


def _data(n=5000, seed=0):
    rng = np.random.RandomState(seed)
    ap = rng.normal(26000, 2500, n); rg = np.abs(rng.normal(5000, 1500, n)); ft = rng.normal(205, 9, n)
    x = rng.normal(4000, 1200, n); y = rng.normal(0, 800, n)
    ap[::17] = rng.choice([np.nan, 4e8, 50.0, 90000.0, -np.inf], ap[::17].size)     # outliers of every kind
    rg[5::91] = 3e5; ft[7::113] = 900.0; rg[11::501] = np.nan
    ap[3] = ap[4]                                                                    # ties
    return ap, rg, ft, x, y


def _check(res, ap, rg, ft, x, y):
    bad = MonteCarloAnalyzer.outlier_mask(ap, rg, ft)
    ok = ~bad
    assert res["n_total"] == ap.size and res["n_samples"] == int(ok.sum()) and res["n_outliers"] == int(bad.sum())
    for key, v in (("apogee_altitude", ap), ("range", rg), ("flight_time", ft)):
        ref = MonteCarloAnalyzer.calc_stats(v[ok])
        for f in ("mean", "std"):
            assert abs(res[key][f] - ref[f]) <= 1e-12 * abs(ref[f])
        assert res[key]["min"] == ref["min"] and res[key]["max"] == ref["max"]
        np.testing.assert_allclose(res[key]["percentiles"], ref["percentiles"], rtol=1e-15, atol=0)
    cov = np.cov(np.vstack([x[ok], y[ok]]), ddof=0)
    np.testing.assert_allclose(res["landing_ellipse"]["covariance"], cov, rtol=1e-11)
    np.testing.assert_allclose(res["landing_ellipse"]["mean"], [x[ok].mean(), y[ok].mean()], rtol=1e-13)


def test_statistics_single_process():
    d = _data()
    res = S.compute_statistics(NumpyBackend(*d), histogram_bins=32)
    _check(res, *d)
    ok = ~MonteCarloAnalyzer.outlier_mask(*d[:3])
    cnt, edges = np.histogram(d[0][ok], bins=32)
    np.testing.assert_array_equal(res["histograms"]["apogee_altitude"]["counts"], cnt)
    assert res["outlier_reasons"]["nonfinite"] > 0 and res["outlier_reasons"]["energy"] > 0


def test_statistics_match_reference_analysis_golden():
    z = util.golden("analysis")
    keep = ~z["in_failed"]
    ap, rg, ft = z["in_apogee"][keep], z["in_range"][keep], z["in_flight_time"][keep]
    res = S.compute_statistics(NumpyBackend(ap, rg, ft, np.zeros_like(ap), np.zeros_like(ap)))
    assert res["n_samples"] == int(z["n_samples"]) and res["n_outliers"] == int(z["n_outliers"])
    for key in ("apogee_altitude", "range", "flight_time"):
        got = np.array([res[key]["mean"], res[key]["std"], res[key]["min"], res[key]["max"], *res[key]["percentiles"]])
        np.testing.assert_allclose(got, z[key], rtol=1e-13)


def test_statistics_edge_cases():
    one = S.compute_statistics(NumpyBackend([25000.0], [100.0], [200.0], [1.0], [2.0]))
    assert one["apogee_altitude"]["percentiles"] == [25000.0] * 5 and one["apogee_altitude"]["std"] == 0.0
    none = S.compute_statistics(NumpyBackend([np.nan, 10.0], [1.0, 1.0], [1.0, 1.0], [0, 0], [0, 0]))
    assert none["n_samples"] == 0 and none["n_outliers"] == 2 and np.isnan(none["range"]["mean"])
    neg = np.array([-5.0, -1.0, 0.0, 3.0, 2.0])            # ordered_keys must sort negatives below positives
    assert np.all(np.diff(S.ordered_keys(np.sort(neg)).astype(np.float64)) > 0)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = _data(6001, seed=3)
    part = [a[rank::world] for a in d]                     # each rank holds an interleaved shard
    res = S.compute_statistics(NumpyBackend(*part, dist=dist), histogram_bins=16)
    q.put((rank, res))
    dist.destroy_process_group()


def test_statistics_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = _data(6001, seed=3)
    for r in (0, 1):
        _check(got[r], *d)                                 # every rank holds the statistics of the WHOLE job
    assert got[0]["apogee_altitude"] == got[1]["apogee_altitude"]
    np.testing.assert_array_equal(got[0]["histograms"]["range"]["counts"], got[1]["histograms"]["range"]["counts"])
