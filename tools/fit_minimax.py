"""Developer tool: Remez minimax fits (mpmath) for the engine's device math (csrc/emc_physics.cuh).
  atan : atan(t)  = t + t*u*Q(u),  u = t^2 in [0, 1]
  log  : atanh(s) = s + s*w*Q(w),  w = s^2 in [0, (3-2*sqrt(2))^2]   (log m = 2 atanh((m-1)/(m+1)), m in [sqrt(.5), sqrt(2)))
  exp  : exp(r)   = 1 + r + r^2*Q(r), r in [-ln2/2, ln2/2]
"""
import sys
import mpmath as mp
mp.mp.dps = 60


def remez(f, w, n, a, b, iters=40, N=4000):
    xs = sorted((a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * (n + 2))) for k in range(n + 2))
    c = None; ev = None
    for it in range(iters):
        A = mp.matrix(n + 2, n + 2); rhs = mp.matrix(n + 2, 1)
        for i, x in enumerate(xs):
            for j in range(n + 1):
                A[i, j] = x ** j
            wi = w(x)
            A[i, n + 1] = ((-1) ** i) / wi if wi != 0 else ((-1) ** i) * mp.mpf(10) ** 40
            rhs[i] = f(x)
        sol = mp.lu_solve(A, rhs)
        c = [sol[j] for j in range(n + 1)]
        err = lambda x: (sum(c[j] * x ** j for j in range(n + 1)) - f(x)) * w(x)
        grid = [a + (b - a) * k / N for k in range(N + 1)]
        ev = [err(g) for g in grid]
        ext = [(grid[0], ev[0])] + [(grid[k], ev[k]) for k in range(1, N) if (ev[k] - ev[k - 1]) * (ev[k + 1] - ev[k]) <= 0] + [(grid[-1], ev[-1])]
        picked = []
        for x, e in ext:
            if not picked or (e > 0) != (picked[-1][1] > 0):
                picked.append((x, e))
            elif abs(e) > abs(picked[-1][1]):
                picked[-1] = (x, e)
        while len(picked) > n + 2:
            picked.pop(0) if abs(picked[0][1]) < abs(picked[-1][1]) else picked.pop()
        if len(picked) < n + 2:
            break
        xs = [p[0] for p in picked]
        mags = [abs(p[1]) for p in picked]
        if max(mags) / min(mags) < 1.0001:
            break
    return c, max(abs(e) for e in ev)


def fit(kind, n):
    if kind == "atan":
        f = lambda u: mp.mpf(-1) / 3 if u == 0 else (mp.atan(mp.sqrt(u)) / mp.sqrt(u) - 1) / u
        w = lambda u: (u / (mp.atan(mp.sqrt(u)) / mp.sqrt(u))) if u > 0 else mp.mpf(0)
        return remez(f, w, n, mp.mpf(0), mp.mpf(1))
    if kind == "log":
        top = (3 - 2 * mp.sqrt(2)) ** 2
        f = lambda v: mp.mpf(1) / 3 if v == 0 else (mp.atanh(mp.sqrt(v)) / mp.sqrt(v) - 1) / v
        w = lambda v: (v / (mp.atanh(mp.sqrt(v)) / mp.sqrt(v))) if v > 0 else mp.mpf(0)
        return remez(f, w, n, mp.mpf(0), top * mp.mpf("1.02"))
    if kind == "exp":
        h = mp.log(2) / 2 * mp.mpf("1.01")
        f = lambda r: mp.mpf(1) / 2 if r == 0 else (mp.exp(r) - 1 - r) / (r * r)
        w = lambda r: (r * r + mp.mpf(10) ** -30) / mp.exp(r)
        return remez(f, w, n, -h, h)
    raise SystemExit("kind?")


if __name__ == "__main__":
    kind = sys.argv[1]
    for n in [int(a) for a in sys.argv[2:]]:
        c, e = fit(kind, n)
        print(f"// {kind}: degree {n}: max relative error {mp.nstr(e, 5)}")
        print("{" + ", ".join(repr(float(x)) for x in c) + "}")
