"""TEST INFRASTRUCTURE: a NumPy stand-in for the statistics kernels (csrc/emc_stats.cuh) with optional gloo
all-reduce, used to exercise the pass sequencing of erpl_monte_carlo_sim_b200/stats.py on CPU (world_size 2)."""
import numpy as np

from erpl_monte_carlo_sim_b200 import stats as S


def classify(ap, rg, ft):
    with np.errstate(invalid="ignore"):
        why = np.zeros(ap.shape, np.int64)
        why |= np.where(~np.isfinite(ap) | ~np.isfinite(rg) | ~np.isfinite(ft), 1, 0)
        why |= np.where(ap > 80000.0, 2, np.where(ap < 100.0, 4, 0))
        why |= np.where(rg > 200000.0, 8, 0) | np.where(ft > 600.0, 16, 0)
        why |= np.where(ap > (1200.0 ** 2 / (2 * 9.81)) * 1.2, 32, 0)
    return why


class NumpyBackend:
    def __init__(self, ap, rg, ft, x, y, dist=None):
        self.v = [np.asarray(a, float) for a in (ap, rg, ft, x, y)]
        self.why = classify(*self.v[:3])
        self.ok = self.why == 0
        self.dist = dist

    def _reduce(self, arr, op="sum"):
        if self.dist is None:
            return arr
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr).copy())
        self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN, "max": self.dist.ReduceOp.MAX}[op])
        return t.numpy()

    def moments1(self):
        ok, why = self.ok, self.why
        s = np.array([why.size, ok.sum(), (~ok).sum()] + [((why & b) != 0).sum() for b in (1, 2, 4, 8, 16, 32)] +
                     [self.v[k][ok].sum() for k in range(5)], float)
        mn = np.array([self.v[k][ok].min() if ok.any() else np.inf for k in range(3)])
        mx = np.array([self.v[k][ok].max() if ok.any() else -np.inf for k in range(3)])
        return self._reduce(s), self._reduce(mn, "min"), self._reduce(mx, "max")

    def moments2(self, c):
        ok = self.ok
        d = [self.v[k][ok] - c[k] for k in range(5)]
        return self._reduce(np.array([np.sum(d[0] ** 2), np.sum(d[1] ** 2), np.sum(d[2] ** 2), np.sum(d[3] ** 2),
                                      np.sum(d[3] * d[4]), np.sum(d[4] ** 2)]))

    def select_hist(self, field, shift, pshift, prefixes):
        key = S.ordered_keys(self.v[field][self.ok])
        pre = np.zeros_like(key) if pshift >= 64 else key >> np.uint64(pshift)
        digit = ((key >> np.uint64(shift)) & np.uint64(S.BINS - 1)).astype(np.int64)
        h = np.zeros((len(prefixes), S.BINS), np.int64)
        for u, p in enumerate(prefixes):
            h[u] = np.bincount(digit[pre == np.uint64(p)], minlength=S.BINS)
        return self._reduce(h)

    def linear_hist(self, field, lo, hi, nbins):
        v = self.v[field][self.ok]
        v = v[(v >= lo) & (v <= hi)]
        b = np.clip(((v - lo) * (nbins / (hi - lo) if hi > lo else 0.0)).astype(np.int64), 0, nbins - 1)
        return self._reduce(np.bincount(b, minlength=nbins).astype(np.int64))
