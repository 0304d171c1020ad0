"""Worker of tests/test_gpu_tape_and_shards.py::test_sharded_run_monte_carlo_equals_single_rank (launched by torchrun):
MonteCarloAnalyzer.run_monte_carlo inside an initialised NCCL process group; every rank dumps its shard and the analysis."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
from erpl_monte_carlo_sim_b200 import MonteCarloAnalyzer, Rocket, SolidMotor, StandardAtmosphere, WindModel  # noqa: E402
from test_host_sampling import CSV_ALT, CSV_WIND  # noqa: E402

out_dir, n = sys.argv[1], int(sys.argv[2])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mc = MonteCarloAnalyzer(Rocket(), SolidMotor(), StandardAtmosphere(), WindModel())
mc.base_altitude_profile, mc.base_wind_profile = CSV_ALT, CSV_WIND
mc.rng = "numpy"
ic = {"position": [0.0, 0.0, 10.0], "velocity": [0, 0, 0.0], "attitude": [0.0, -np.pi / 2 + 0.02, 0.0], "angular_velocity": [0.0, 0.0, 0.0]}
an = mc.run_monte_carlo(ic, n_samples=n)
run = mc.last_run
keep = {k: an[k] for k in ("n_samples", "n_outliers", "n_failed", "apogee_altitude", "range", "flight_time", "parameter_ranges_observed", "shard")}
np.savez(os.path.join(out_dir, f"rank{dist.get_rank()}.npz"), out=run.out, iout=run.iout, first_id=run.first_id,
         analysis=json.dumps(keep))
dist.barrier()
dist.destroy_process_group()
