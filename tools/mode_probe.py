"""developer tool (GPU box): where does run_monte_carlo(ic, 100_000) spend its wall time in each RNG mode?"""
import gc, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from erpl_monte_carlo_sim_b200.simulator import get_engine

for mode in ("numpy-device", "philox"):
    mc = bench.c3_analyzer()
    mc.rng = mode; mc.host_rng_max = 1 << 40
    for rep in range(4):
        mc.last_run = None; gc.collect()
        t0 = time.perf_counter()
        an = mc.run_monte_carlo(bench.IC_C3, n_samples=100_000)
        dt = time.perf_counter() - t0
        c = get_engine(0).counters()
        it = mc.last_run.iout if rep == 3 else None
        print(mode, rep, "wall %.1f ms" % (dt * 1e3), "flight %.1f ms" % c["flight_ms"], "rail %.2f" % c["rail_ms"], "strict %.2f" % c["strict_ms"],
              "rk4 steps", c["rk4_steps"], "parked", c["parked"], "longest flight (steps)", int(it[0].max()) if it is not None else "")
        an = None
