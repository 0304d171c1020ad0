"""Developer tool: minimax coefficients for atan(t) = t + t*u*Q(u), u = t*t in [0,1] (mpmath Remez).
Prints the double-precision coefficients used by fast_atan2 in csrc/emc_physics.cuh."""
import sys
import mpmath as mp
mp.mp.dps = 60

def f(u):
    if u == 0:
        return mp.mpf(-1) / 3
    t = mp.sqrt(u)
    return (mp.atan(t) / t - 1) / u          # Q(u)

def remez(n, a=mp.mpf(0), b=mp.mpf(1), iters=30):
    # weight: absolute error of Q times u*t relative to atan(t) ~ t -> relative error of atan = u*|dQ|/(atan/t)
    w = lambda u: (u / (mp.atan(mp.sqrt(u)) / mp.sqrt(u))) if u > 0 else mp.mpf(0)
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * (n + 2))) for k in range(n + 2)]
    xs = sorted(xs)
    for it in range(iters):
        A = mp.matrix(n + 2, n + 2); rhs = mp.matrix(n + 2, 1)
        for i, x in enumerate(xs):
            for j in range(n + 1):
                A[i, j] = x ** j
            wi = w(x)
            A[i, n + 1] = ((-1) ** i) / wi if wi != 0 else ((-1) ** i) * mp.mpf(10) ** 30
            rhs[i] = f(x)
        sol = mp.lu_solve(A, rhs)
        c = [sol[j] for j in range(n + 1)]
        err = lambda x: (sum(c[j] * x ** j for j in range(n + 1)) - f(x)) * w(x)
        # find extrema on a fine grid
        N = 4000
        grid = [a + (b - a) * k / N for k in range(N + 1)]
        ev = [err(g) for g in grid]
        ext = []
        for k in range(1, N):
            if (ev[k] - ev[k - 1]) * (ev[k + 1] - ev[k]) <= 0:
                ext.append((grid[k], ev[k]))
        ext = [(grid[0], ev[0])] + ext + [(grid[-1], ev[-1])]
        # pick alternating extrema with largest magnitude
        picked = []
        for x, e in ext:
            if not picked or (e > 0) != (picked[-1][1] > 0):
                picked.append((x, e))
            elif abs(e) > abs(picked[-1][1]):
                picked[-1] = (x, e)
        while len(picked) > n + 2:
            if abs(picked[0][1]) < abs(picked[-1][1]):
                picked.pop(0)
            else:
                picked.pop()
        if len(picked) < n + 2:
            break
        xs = [p[0] for p in picked]
        emax = max(abs(p[1]) for p in picked); emin = min(abs(p[1]) for p in picked)
        if emax / emin < 1.0001:
            break
    return c, max(abs(e) for e in ev)

if __name__ == "__main__":
    import os
    umax = mp.mpf(os.environ.get("ATAN_UMAX", "1"))          # 0.1715728752538099 = tan(pi/8)^2 for the two-step reduction
    for n in [int(a) for a in sys.argv[1:]] or [16, 18, 20]:
        c, e = remez(n, b=umax)
        print(f"// degree {n} in u: max weighted (relative) error {mp.nstr(e, 5)}")
        print("{" + ", ".join(repr(float(x)) for x in c) + "}")
