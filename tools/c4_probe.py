"""Developer tool (GPU box): config C4 through the API at a small size, printing the engine's error text if any."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from erpl_monte_carlo_sim_b200 import LiquidMotor, MonteCarloAnalyzer, Rocket, StandardAtmosphere, WindModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
mc = MonteCarloAnalyzer(Rocket(), LiquidMotor(), StandardAtmosphere(), WindModel())
mc.rng = "philox"; mc.trajectory_samples = 0
if len(sys.argv) > 2 and "opts" in sys.argv[2]:
    from erpl_monte_carlo_sim_b200 import _lib
    mc.run_opts = _lib.run_opts(refill_threshold=0, block_threads=0, blocks_per_sm=0, cold_state_in_smem=2)
if len(sys.argv) > 2 and "torch" in sys.argv[2]:
    import torch
    x = torch.zeros(1 << 20, device="cuda:0"); torch.cuda.synchronize()
if len(sys.argv) > 2 and "eng" in sys.argv[2]:
    from erpl_monte_carlo_sim_b200 import _lib
    e1 = _lib.Engine(0)
    md, blk, wind, _ = bench.make_workload("c3", 4096, 0)
    e1.set_model(md); e1.run_batch(blk, wind); print("first engine ran", e1.counters()["parked"], flush=True)
for rep in range(2):
    t0 = time.perf_counter()
    try:
        an = mc.run_monte_carlo(bench.IC_C4, n_samples=n)
        print("ok", n, time.perf_counter() - t0, an["n_samples"], bench.eng_counters_of(mc), flush=True)
    except Exception as e:
        print("FAILED", repr(e)[:600], flush=True)
        break
