"""Device-side Monte Carlo statistics (reference monte_carlo.py:337-473) driven from the host.

The kernels (csrc/emc_stats.cuh) classify outliers and reduce counts / sums / min / max / centred second
moments / digit histograms on the GPU.  This module sequences the passes and, in a multi-GPU job, all-reduces the
small result blocks between them with `torch.distributed` (NCCL over NVLink): samples never leave their GPU, yet
mean, std, min, max, the landing-ellipse covariance and np.percentile's order statistics are exact over the job.

`compute_statistics` is written against a small backend interface so that the pass sequencing (and its
multi-rank reduction) can be exercised on CPU with world_size-2 `gloo` in tests/ (NumpyBackend); the product
path always uses DeviceBackend.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

SUM_FIELDS = ["n", "valid", "outlier", "nonfinite", "apogee_high", "apogee_low", "range", "time", "energy",
              "sum_apogee", "sum_range", "sum_time", "sum_x", "sum_y"]
NSUM, NMM, N2, BINS, MAXP = len(SUM_FIELDS), 3, 6, 2048, 16
METRICS = ("apogee_altitude", "range", "flight_time")
PERCENTILES = (5, 25, 50, 75, 95)
# radix-select digits of the order-preserving 64-bit key: 9 + 5 x 11 bits -> (shift, prefix_shift)
SELECT_PASSES = [(55, 64), (44, 55), (33, 44), (22, 33), (11, 22), (0, 11)]


class DeviceBlock:
    """A typed view of engine-owned device memory; exposes __cuda_array_interface__ so that
    torch.as_tensor(block, device=...) aliases it without a copy (for dist.all_reduce over NCCL)."""

    def __init__(self, engine, ptr, count, dtype):
        self.engine, self.ptr, self.count, self.dtype = engine, int(ptr), int(count), np.dtype(dtype)
        self.__cuda_array_interface__ = {"shape": (self.count,), "typestr": self.dtype.str, "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    @property
    def nbytes(self):
        return self.count * self.dtype.itemsize

    def get(self):
        host = np.empty(self.count, self.dtype)
        e = self.engine
        e._check(e._lib.emc_copy_to_host(e._ctx, host.ctypes.data, C.c_void_p(self.ptr), self.nbytes), "emc_copy_to_host")
        return host

    def put(self, host):
        host = np.ascontiguousarray(host, self.dtype)
        assert host.size == self.count
        e = self.engine
        e._check(e._lib.emc_copy_to_device(e._ctx, C.c_void_p(self.ptr), host.ctypes.data, self.nbytes), "emc_copy_to_device")


def _dist_active():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


class DeviceBackend:
    """Statistics passes on the GPU (csrc/emc_stats.cuh); reductions across ranks on device memory (NCCL)."""

    def __init__(self, engine, n, out_dev=None, ld=None, distributed=None):
        self.e, self.n = engine, int(n)
        self.out_ptr = C.c_void_p(out_dev) if out_dev else None
        self.ld = int(ld) if ld is not None else 0
        self.distributed = _dist_active() if distributed is None else bool(distributed)
        self.stream_ordered_collectives = True       # NCCL all-reduce can be enqueued on the engine's stream
        words = NSUM + 2 * NMM + 5 + N2 + 3 * MAXP * BINS
        base = C.c_void_p()
        engine._check(engine._lib.emc_scratch(engine._ctx, words * 8, C.byref(base)), "emc_scratch")
        self.base = base.value

    def _blk(self, off, cnt, dt=np.float64):
        return DeviceBlock(self.e, self.base + 8 * off, cnt, dt)

    def _reduce(self, block, op="sum"):
        if not self.distributed:
            return
        import torch
        import torch.distributed as dist
        t = torch.as_tensor(block, device=torch.device("cuda", self.e.device))
        if block.dtype == np.uint64:
            t = t.view(torch.int64)                       # counts fit int64
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op])
        torch.cuda.synchronize(self.e.device)

    def moments1(self):
        e, vp = self.e, C.c_void_p
        bs, bmin, bmax = self._blk(0, NSUM), self._blk(NSUM, NMM), self._blk(NSUM + NMM, NMM)
        e._check(e._lib.emc_stats_moments1(e._ctx, self.out_ptr, self.ld, self.n, vp(bs.ptr), vp(bmin.ptr), vp(bmax.ptr)),
                 "emc_stats_moments1")
        self._reduce(bs, "sum"); self._reduce(bmin, "min"); self._reduce(bmax, "max")
        allv = self._blk(0, NSUM + 2 * NMM).get()                 # one device->host copy for the three blocks
        return allv[:NSUM], allv[NSUM:NSUM + NMM], allv[NSUM + NMM:]

    def summary(self, percentiles):
        """Every pass chained on the device, one host synchronisation.  One GPU: emc_stats_summary.  Several GPUs: the same
        chain stage by stage (emc_stats_summary_stage) with the small blocks all-reduced between stages ON THE ENGINE'S
        STREAM (NCCL, stream-ordered: no host round trip either).
        Returns (sum, min, max, s2, val) with val[f][2j], val[f][2j+1] the lo / hi order statistic of percentile j."""
        e = self.e
        npct = len(percentiles)
        pct = (C.c_double * npct)(*[float(p) for p in percentiles])
        res = np.empty(32 + 3 * 2 * npct, np.float64)
        rp = res.ctypes.data_as(C.POINTER(C.c_double))
        if self.distributed:
            import torch
            import torch.distributed as dist
            sp = C.c_void_p()
            e._check(e._lib.emc_stream(e._ctx, C.byref(sp)), "emc_stream")
            dev = torch.device("cuda", e.device)
            ext = torch.cuda.ExternalStream(sp.value, device=dev)
            with torch.cuda.stream(ext):
                for stage in range(15):
                    bp, bw = C.c_void_p(), C.c_int64()
                    e._check(e._lib.emc_stats_summary_stage(e._ctx, self.out_ptr, self.ld, self.n, pct, npct, stage, C.byref(bp),
                                                            C.byref(bw), rp), "emc_stats_summary_stage")
                    if bp.value:
                        if stage == 0:
                            for off, cnt, op in ((0, NSUM, dist.ReduceOp.SUM), (NSUM, NMM, dist.ReduceOp.MIN), (NSUM + NMM, NMM, dist.ReduceOp.MAX)):
                                dist.all_reduce(torch.as_tensor(DeviceBlock(e, bp.value + 8 * off, cnt, np.float64), device=dev), op=op)
                        elif stage == 1:
                            dist.all_reduce(torch.as_tensor(DeviceBlock(e, bp.value, bw.value, np.float64), device=dev))
                        else:
                            dist.all_reduce(torch.as_tensor(DeviceBlock(e, bp.value, bw.value, np.uint64), device=dev).view(torch.int64))
        else:
            e._check(e._lib.emc_stats_summary(e._ctx, self.out_ptr, self.ld, self.n, pct, npct, rp), "emc_stats_summary")
        return (res[:NSUM], res[NSUM:NSUM + NMM], res[NSUM + NMM:NSUM + 2 * NMM], res[20:26],
                res[32:].reshape(3, 2 * npct))

    def select_hist3(self, shift, prefix_shift, prefix_lists):
        """One pass over the samples for all three metrics: prefix_lists[f] = list of prefixes of metric f."""
        e = self.e
        pre = (C.c_uint64 * (3 * MAXP))()
        cnt = (C.c_int32 * 3)()
        for f in range(3):
            cnt[f] = len(prefix_lists[f])
            for u, p in enumerate(prefix_lists[f]):
                pre[f * MAXP + u] = p
        rows = sum(len(p) for p in prefix_lists)
        bh = self._blk(NSUM + 2 * NMM + 5 + N2, rows * BINS, np.uint64)          # compact: only the rows in use
        e._check(e._lib.emc_stats_select_hist3(e._ctx, self.out_ptr, self.ld, self.n, shift, prefix_shift, pre, cnt,
                                               C.c_void_p(bh.ptr)), "emc_stats_select_hist3")
        self._reduce(bh, "sum")
        h = bh.get().view(np.int64).reshape(rows, BINS)
        edges = np.cumsum([0] + [len(p) for p in prefix_lists])
        return [h[edges[f]:edges[f + 1]] for f in range(3)]

    def moments2(self, center):
        e, vp = self.e, C.c_void_p
        bc, b2 = self._blk(NSUM + 2 * NMM, 5), self._blk(NSUM + 2 * NMM + 5, N2)
        bc.put(center)
        e._check(e._lib.emc_stats_moments2(e._ctx, self.out_ptr, self.ld, self.n, vp(bc.ptr), vp(b2.ptr)), "emc_stats_moments2")
        self._reduce(b2, "sum")
        return b2.get()

    def select_hist(self, field, shift, prefix_shift, prefixes):
        e = self.e
        bh = self._blk(NSUM + 2 * NMM + 5 + N2, len(prefixes) * BINS, np.uint64)
        pre = (C.c_uint64 * len(prefixes))(*prefixes)
        e._check(e._lib.emc_stats_select_hist(e._ctx, self.out_ptr, self.ld, self.n, field, shift, prefix_shift, pre,
                                              len(prefixes), C.c_void_p(bh.ptr)), "emc_stats_select_hist")
        self._reduce(bh, "sum")
        return bh.get().reshape(len(prefixes), BINS).astype(np.int64)

    def linear_hist(self, field, lo, hi, nbins):
        e = self.e
        bh = self._blk(NSUM + 2 * NMM + 5 + N2, nbins, np.uint64)
        e._check(e._lib.emc_stats_linear_hist(e._ctx, self.out_ptr, self.ld, self.n, field, C.c_double(lo), C.c_double(hi),
                                              nbins, C.c_void_p(bh.ptr)), "emc_stats_linear_hist")
        self._reduce(bh, "sum")
        return bh.get().astype(np.int64)


def allreduce_minmax(engine, lo, hi):
    """Element-wise min of `lo` and max of `hi` over the ranks of the initialised process group (NCCL on the engine's GPU,
    gloo on the host)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", engine.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    a, b = torch.as_tensor(np.ascontiguousarray(lo), device=dev), torch.as_tensor(np.ascontiguousarray(hi), device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.MIN); dist.all_reduce(b, op=dist.ReduceOp.MAX)
    return a.cpu().numpy(), b.cpu().numpy()


def ordered_keys(v):
    """Order-preserving map float64 -> uint64 (the device kernel's `ordered_key`)."""
    b = np.asarray(v, np.float64).view(np.uint64)
    return np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))


def _key_to_double(key: int) -> float:
    b = (key & 0x7FFFFFFFFFFFFFFF) if (key >> 63) else (~key & 0xFFFFFFFFFFFFFFFF)
    return float(np.array([b], np.uint64).view(np.float64)[0])


def _lerp(a, b, t):
    """numpy/lib/_function_base_impl.py::_lerp — the interpolation np.percentile(method='linear') applies."""
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def radix_select3(backend, ranks):
    """radix_select for the three metrics in lockstep (one pass over the samples per digit) when the backend offers
    `select_hist3` and every metric needs <= MAXP distinct prefixes; otherwise metric by metric."""
    if not hasattr(backend, "select_hist3") or len(ranks) > MAXP:
        return [radix_select(backend, f, ranks) for f in range(3)]
    targets = [{r: (0, r) for r in ranks} for _ in range(3)]
    for shift, pshift in SELECT_PASSES:
        width = (64 - shift) if pshift >= 64 else (pshift - shift)
        uniq = [sorted({t[0] for t in targets[f].values()}) for f in range(3)]
        hists = backend.select_hist3(shift, pshift, uniq)
        for f in range(3):
            cums = {pre: np.cumsum(hists[f][u]) for u, pre in enumerate(uniq[f])}
            nxt = {}
            for r, (pre, rem) in targets[f].items():
                cum = cums[pre]
                b = int(np.searchsorted(cum, rem, side="right"))
                nxt[r] = ((pre << width) | b, rem - (int(cum[b - 1]) if b > 0 else 0))
            targets[f] = nxt
    return [{r: _key_to_double(t[0]) for r, t in targets[f].items()} for f in range(3)]


def radix_select(backend, field, ranks):
    """Exact order statistics (0-based global ranks) of metric `field` over all valid samples of all ranks."""
    targets = {r: (0, r) for r in ranks}                  # rank -> (prefix so far, remaining rank inside it)
    for shift, pshift in SELECT_PASSES:
        width = (64 - shift) if pshift >= 64 else (pshift - shift)
        uniq = sorted({t[0] for t in targets.values()})
        cums = {}
        for c0 in range(0, len(uniq), MAXP):
            chunk = uniq[c0:c0 + MAXP]
            h = backend.select_hist(field, shift, pshift, chunk)
            for u, pre in enumerate(chunk):
                cums[pre] = np.cumsum(h[u])
        nxt = {}
        for r, (pre, rem) in targets.items():
            cum = cums[pre]
            b = int(np.searchsorted(cum, rem, side="right"))
            nxt[r] = ((pre << width) | b, rem - (int(cum[b - 1]) if b > 0 else 0))
        targets = nxt
    return {r: _key_to_double(t[0]) for r, t in targets.items()}


def compute_statistics(backend, histogram_bins=0):
    fused = None
    if hasattr(backend, "summary") and getattr(backend, "fused", True) and len(PERCENTILES) <= 8 \
            and (not getattr(backend, "distributed", True) or getattr(backend, "stream_ordered_collectives", False)):
        fused = backend.summary(PERCENTILES)          # one stream-ordered chain on the device, one host sync
        s, mn, mx = fused[0], fused[1], fused[2]
    else:
        s, mn, mx = backend.moments1()
    S = dict(zip(SUM_FIELDS, s))
    m = int(round(S["valid"]))
    res = {"n_total": int(round(S["n"])), "n_samples": m, "n_outliers": int(round(S["outlier"])), "n_failed": 0,
           "outlier_reasons": {k: int(round(S[k])) for k in ("nonfinite", "apogee_high", "apogee_low", "range", "time", "energy")}}
    nan = float("nan")
    if m == 0:
        for k in METRICS:
            res[k] = {"mean": nan, "std": nan, "min": nan, "max": nan, "percentiles": [nan] * 5}
        res["landing_ellipse"] = {"mean": [nan, nan], "covariance": [[nan, nan], [nan, nan]]}
        return res
    means = np.array([S["sum_apogee"], S["sum_range"], S["sum_time"], S["sum_x"], S["sum_y"]]) / m
    s2 = fused[3] if fused is not None else backend.moments2(means)
    for f, key in enumerate(METRICS):
        res[key] = {"mean": float(means[f]), "std": float(math.sqrt(s2[f] / m)), "min": float(mn[f]), "max": float(mx[f])}
    res["landing_ellipse"] = {"mean": [float(means[3]), float(means[4])],
                              "covariance": [[float(s2[3] / m), float(s2[4] / m)], [float(s2[4] / m), float(s2[5] / m)]]}
    # exact np.percentile(values, [5, 25, 50, 75, 95]) via distributed radix select of the bracketing order statistics
    q = np.true_divide(np.asarray(PERCENTILES, np.float64), 100)
    pos = (m - 1) * q
    lo_idx = np.floor(pos).astype(np.int64)
    hi_idx = np.minimum(lo_idx + 1, m - 1)
    gamma = pos - lo_idx
    ranks = sorted(set(lo_idx.tolist()) | set(hi_idx.tolist()))
    if fused is not None:
        for f, key in enumerate(METRICS):
            v = fused[4][f]
            res[key]["percentiles"] = [float(_lerp(v[2 * j], v[2 * j + 1], g)) for j, g in enumerate(gamma)]
    else:
        vals = radix_select3(backend, ranks)
        for f, key in enumerate(METRICS):
            val = vals[f]
            res[key]["percentiles"] = [float(_lerp(val[int(a)], val[int(b)], g)) for a, b, g in zip(lo_idx, hi_idx, gamma)]
    if histogram_bins:
        res["histograms"] = {}
        for f, key in enumerate(METRICS + ("landing_x", "landing_y")):
            if f < 3:
                lo, hi = float(mn[f]), float(mx[f])
            else:
                c, sd = means[f], math.sqrt(s2[3 if f == 3 else 5] / m)
                lo, hi = float(c - 4 * sd), float(c + 4 * sd)
            res["histograms"][key] = {"edges": np.linspace(lo, hi, histogram_bins + 1),
                                      "counts": backend.linear_hist(f, lo, hi, histogram_bins)}
    return res


def device_statistics(engine, n, out_dev=None, ld=None, distributed=None, histogram_bins=0, fused=True):
    """Statistics of the `n` samples in `out_dev` (device pointer; None = the engine's last run_batch outputs).
    fused=False forces the pass-by-pass path a multi-GPU job uses (blocks all-reduced between passes)."""
    backend = DeviceBackend(engine, n, out_dev, ld, distributed)
    backend.fused = bool(fused)
    return compute_statistics(backend, histogram_bins)
