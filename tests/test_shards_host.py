"""CPU: host logic of the round-2 boundary — shard ranges, the lazy per-sample result dict and the trajectory store of a
BatchRun, min/max reduction of the observed parameter ranges over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from erpl_monte_carlo_sim_b200 import _abi, stats
from erpl_monte_carlo_sim_b200.monte_carlo import BatchRun, DispersionSet, SampleDict, SampleResults, shard_range


def test_shard_ranges_tile_the_sample_index_range():
    for n in (0, 1, 7, 100, 100_000, 10_000_000, 12345679):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


class _FakeAnalyzer:
    def __init__(self):
        self.taped = []

    def _tape_samples(self, run, ids):
        self.taped.append(list(ids))
        rows = np.zeros((len(ids), 3, _abi.BTAPE_WIDTH)); rows[:, :, 0] = [0.0, 0.1, 0.17]
        rows[:, :, 3] = np.asarray(ids, float)[:, None]
        run.add_tape(ids, rows, np.full(len(ids), 3, np.int32), 20)


def test_sample_dict_serves_the_trajectory_lazily():
    n = 6
    out = np.arange(_abi.OUT_COUNT * n, dtype=float).reshape(_abi.OUT_COUNT, n)
    iout = np.ones((_abi.IOUT_COUNT, n), np.int32)
    fa = _FakeAnalyzer()
    run = BatchRun(fa, {}, DispersionSet(n), out, iout, None, np.zeros((_abi.IN_COUNT, n)), first_id=100)
    rows = np.full((2, 5, _abi.BTAPE_WIDTH), np.nan)
    rows[0, :4] = [[0, 1, 2, 3], [0.1, 4, 5, 6], [0.2, 7, 8, 9], [0.25, 10, 11, 12]]
    rows[1, :2] = [[0, 0, 0, 0], [0.005, 1, 1, 1]]
    run.add_tape([0, 4], rows, [4, 2], 20)
    res = SampleResults(run, [4, 5, 0])
    d = res[0]
    assert isinstance(d, SampleDict) and d["simulation_id"] == 104
    assert "trajectory" in d and "trajectory" not in dict.keys(d)              # promised, not yet materialised
    tr = d["trajectory"]
    assert tr["time"].tolist() == [0.0, 0.005] and tr["position"].shape == (2, 3) and tr["altitude"].tolist() == [0.0, 1.0]
    assert "trajectory" in dict.keys(d) and fa.taped == []                      # served from the recorded tape
    e = res[1]                                                                  # sample 5 was not taped: one extra batch
    assert e.get("trajectory")["altitude"].tolist() == [5.0, 5.0, 5.0] and fa.taped == [[5]]
    run.ensure_trajectories([0, 4, 5, 1, 2])
    assert fa.taped == [[5], [1, 2]]
    assert d.get("missing", 7) == 7
    full = res[2]["trajectory"]
    assert full["time"].tolist() == [0.0, 0.1, 0.2, 0.25] and full["position"][3].tolist() == [10.0, 11.0, 12.0]


def _worker(rank, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    lo = np.array([1.0, -3.0, np.inf]) if rank == 0 else np.array([0.5, 2.0, 4.0])
    hi = np.array([1.0, -3.0, -np.inf]) if rank == 0 else np.array([0.5, 2.0, 4.0])
    a, b = stats.allreduce_minmax(None, lo, hi)
    ret[rank] = (a.tolist(), b.tolist())
    dist.destroy_process_group()


def test_parameter_ranges_reduce_over_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    ps = [ctx.Process(target=_worker, args=(r, port, ret)) for r in range(2)]
    [p.start() for p in ps]; [p.join(120) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for r in range(2):
        assert ret[r] == ([0.5, -3.0, 4.0], [1.0, 2.0, 4.0])      # an empty shard (+inf / -inf) does not disturb the result


def test_lazy_analysis_keys_behave_like_the_reference_dict():
    from erpl_monte_carlo_sim_b200.monte_carlo import LazyAnalysis
    calls = []
    a = LazyAnalysis({"n_samples": 3}, {"parameter_ranges_observed": lambda: (calls.append(1), {"mass_multiplier": {"min": 0.9, "max": 1.1}})[1]})
    assert "parameter_ranges_observed" in a and calls == []                    # promised, not computed
    assert a.get("nope", 7) == 7 and calls == []
    assert set(a) == {"n_samples", "parameter_ranges_observed"} and calls == [1]   # enumeration materialises it, once
    assert a["parameter_ranges_observed"]["mass_multiplier"]["max"] == 1.1 and len(a) == 2 and calls == [1]
