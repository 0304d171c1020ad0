"""Host-side Philox4x32-10 (Salmon et al., Random123) mirroring csrc/emc_philox.cuh: used to reconstruct the
parameters of device-generated samples and by the tests (known-answer vectors, bit-exact uniforms)."""
from __future__ import annotations

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over counters (uint64 arrays holding 32-bit words) -> four uint64 arrays of 32-bit words."""
    c = [np.asarray(x, np.uint64) for x in (c0, c1, c2, c3)]
    k0 = np.uint64(k0); k1 = np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]; p1 = np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & _MASK, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & _MASK]
        k0 = (k0 + np.uint64(W0)) & _MASK; k1 = (k1 + np.uint64(W1)) & _MASK
    return c


def _u01(hi, lo):
    b = ((hi << np.uint64(32)) | lo) >> np.uint64(11)
    return (b.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def uniforms(seed, idx):
    """The two uniforms of sample(s) idx (stream 1, block 0) -> (n, 2)."""
    idx = np.atleast_1d(np.asarray(idx, np.uint64))
    r = philox4x32_10(idx & _MASK, idx >> np.uint64(32), np.zeros_like(idx), np.ones_like(idx), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack([_u01(r[0], r[1]), _u01(r[2], r[3])], axis=1)


def normals(seed, idx, n_gauss):
    """The first n_gauss standard normals of sample(s) idx (stream 0; Box-Muller per Philox block) -> (n, n_gauss)."""
    idx = np.atleast_1d(np.asarray(idx, np.uint64))
    nb = (n_gauss + 1) // 2
    out = np.empty((idx.size, 2 * nb))
    for j in range(nb):
        r = philox4x32_10(idx & _MASK, idx >> np.uint64(32), np.full_like(idx, j), np.zeros_like(idx), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        u1, u2 = _u01(r[0], r[1]), _u01(r[2], r[3])
        rad = np.sqrt(-2.0 * np.log(u1))
        out[:, 2 * j] = rad * np.cos(2.0 * np.pi * u2)
        out[:, 2 * j + 1] = rad * np.sin(2.0 * np.pi * u2)
    return out[:, :n_gauss]
