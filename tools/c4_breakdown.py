"""Developer tool (GPU box): where the wall time of config C4 through the API goes (1.25 M samples, one GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from erpl_monte_carlo_sim_b200 import LiquidMotor, MonteCarloAnalyzer, Rocket, StandardAtmosphere, WindModel, marshal, stats
from erpl_monte_carlo_sim_b200.simulator import get_engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1250000
mc = MonteCarloAnalyzer(Rocket(), LiquidMotor(), StandardAtmosphere(), WindModel())
mc.rng = "philox"; mc.trajectory_samples = 0
eng = get_engine(0)
for rep in range(3):
    t = [time.perf_counter()]
    alts = mc._altitude_grid()
    eng.set_model(marshal.model_dict(mc.rocket, mc.motor, mc.atmosphere, mc._model_simulator(), alts)); t.append(time.perf_counter())
    ds = mc.dispersion_struct(bench.IC_C4); t.append(time.perf_counter())
    eng.generate_inputs(ds, mc.philox_seed, 0, n); t.append(time.perf_counter())
    eng.run_batch_staged(n, download=False); t.append(time.perf_counter())
    c = eng.counters()
    st = stats.device_statistics(eng, n, histogram_bins=mc.histogram_bins); t.append(time.perf_counter())
    names = ["set_model", "dispersion_struct", "generate_inputs", "run_batch_staged", "device_statistics"]
    print({k: round((b - a) * 1e3, 2) for k, a, b in zip(names, t[:-1], t[1:])}, "flight_ms", round(c["flight_ms"], 2), "rail_ms", round(c["rail_ms"], 2), flush=True)
t0 = time.perf_counter(); an = mc.run_monte_carlo(bench.IC_C4, n_samples=n); print("run_monte_carlo", round((time.perf_counter() - t0) * 1e3, 2), "ms")
