"""Developer probe (GPU box): where MonteCarloAnalyzer.run_monte_carlo spends its wall time (cProfile, cumulative)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from erpl_monte_carlo_sim_b200 import LiquidMotor, MonteCarloAnalyzer, Rocket, StandardAtmosphere, WindModel  # noqa: E402

cases = [("c3 100k numpy-device", bench.c3_analyzer(), "numpy-device", bench.IC_C3, 100_000),
         ("c3 100k philox", bench.c3_analyzer(), "philox", bench.IC_C3, 100_000),
         ("c4 1.25M philox", MonteCarloAnalyzer(Rocket(), LiquidMotor(), StandardAtmosphere(), WindModel()), "philox", bench.IC_C4, 1_250_000)]
for name, mc, mode, ic, n in cases:
    mc.rng = mode
    mc.run_monte_carlo(ic, n_samples=min(n, 100_000))          # warm-up (allocations, first launch)
    t0 = time.perf_counter(); mc.run_monte_carlo(ic, n_samples=n); dt = time.perf_counter() - t0
    pr = cProfile.Profile(); pr.enable(); mc.run_monte_carlo(ic, n_samples=n); pr.disable()
    print(f"=== {name}: {dt * 1e3:.1f} ms wall, {n / dt:.0f} traj/s")
    st = pstats.Stats(pr); st.sort_stats("cumulative")
    st.print_stats(14)
