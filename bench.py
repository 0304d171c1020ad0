"""bench.py — headline benchmark of the hot path (BASELINE.json): Monte Carlo trajectories/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|planar]

A "step" is one pass of the hot path (rail phase + RK4 flight + summaries) over one batch of dispersed
samples.  Default workload = BASELINE config C3: SolidMotor + sample_wind.csv altitude-resolved wind +
stochastic perturbations, 100 000 host-seeded samples per GPU (weak scaling; rank r flies seeds
r*S .. (r+1)*S-1).  One JSON line on stdout (rank 0).

  value   trajectories/s with the inputs resident in HBM (device-pointer entry of the C ABI)
  e2e     same metric through the host-buffer entry emc_run_batch, pinned host buffers, H2D + D2H inside
  roofline.bound = "fp64": achieved = RK4 steps/s x 1600 flop (SURVEY.md §8d canonical count) against the
          DFMA peak measured in the same run (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline   the C oracle (port of the reference path) on this box's host cores, bounded sample
  --impl reference   the same CPU port timed as its own arm (the Python reference cannot travel to the box)
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of ONE flight-kernel launch on this workload (100 k C3 samples), from
# `ncu --set full` (profiles/r1_flight_kernel_bench.txt): 32.28 MB + 0.54 MB.  Algorithmic bytes: 30.4 MB in + 30 MB out.
FLIGHT_KERNEL_DRAM_BYTES_100K = 32.82e6
STATS_LAUNCHES_FUSED = 2 + 1 + 2 + 12   # the statistics chain: moments1 + finish, plan, moments2 + finish, 6 x (digit histogram + digit decision); NCCL kernels not counted
FLOP_PER_STEP = 1600.0          # 4*348 + 203 ~ 1.6 kflop per accepted RK4 step
CSV_ALT = np.array([0.0, 5000.0, 10000.0, 15000.0, 20000.0, 25000.0])                # sample_wind.csv:2-7
CSV_WIND = np.array([[2.0, 0, 0], [5, 1, 0], [8, 2, 0], [10, 2, 0], [12, 3, 0], [15, 3, 0]], float)


def make_workload(name, n, first_seed):
    """Host-seeded dispersions of the named configuration -> (model dict, scalars, wind, description)."""
    from erpl_monte_carlo_sim_b200 import (MonteCarloAnalyzer, Rocket, SolidMotor, StandardAtmosphere, WindModel, marshal)
    mc = MonteCarloAnalyzer(Rocket(), SolidMotor(), StandardAtmosphere(), WindModel())
    mc.base_altitude_profile, mc.base_wind_profile = CSV_ALT, CSV_WIND.copy()
    ic = {"position": [0.0, 0.0, 10.0], "velocity": [0.0, 0.0, 0.0], "attitude": [0.0, -np.pi / 2 + 0.02, 0.0],
          "angular_velocity": [0.0, 0.0, 0.0]}
    disp = mc.draw_parameters(n, first_seed=first_seed)
    desc = "C3: SolidMotor + sample_wind.csv 6-knot wind + stochastic perturbations, reference default dispersions, vertical launch"
    if name == "planar":
        # W-B: the same dispersions projected onto the pitch plane so flights reach landing
        disp.pos[:, 1] = 0.0; disp.vel[:, 1] = 0.0; disp.att[:, 0] = 0.0; disp.att[:, 2] = 0.0
        disp.omega[:, 0] = 0.0; disp.omega[:, 2] = 0.0
        disp.wind_direction[:] = np.where(disp.seed % 2 == 0, 0.0, np.pi)
        desc = "W-B planar projection of C3 (beta == 0): launch -> apogee -> parachute -> landing"
    blk, wind, alts = mc.build_inputs(ic, disp)
    if name == "planar":
        wind[:, :, 1] = 0.0
    md = marshal.model_dict(mc.rocket, mc.motor, mc.atmosphere, mc._model_simulator(), alts)
    return md, blk, wind, desc


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_baseline(md, blk, wind, n_cpu, threads=0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib as O
    n_cpu = min(n_cpu, blk.shape[1])
    sc = np.ascontiguousarray(blk[:, :n_cpu]); w = np.ascontiguousarray(wind[:n_cpu])
    cores = O.max_threads() if threads <= 0 else threads
    t0 = time.perf_counter()
    out, iout = O.batch(md, sc, w, n_threads=cores, diagnostics=True)
    dt = time.perf_counter() - t0
    return {"value": n_cpu / dt, "unit": "trajectories/s", "cores": cores, "kind": "port",
            "sample": f"first {n_cpu} samples of rank 0's batch, oracle/emc_oracle.c (C port of the reference path, "
                      f"pthreads), {int(iout[0].sum())} RK4 steps in {dt:.2f} s",
            "steps_per_s": float(iout[0].sum() / dt)}, dt


def secondary_measurements(eng, opts, peak_tf):
    """Not the headline: the same kernel (a) on a batch large enough that the drain of the work queue is amortised
    (C4's per-GPU share is 1.25 M samples) and (b) on the planar W-B set whose flights really reach landing."""
    out = {}
    for key, workload, n in (("c3_1M_samples", "c3", 1_000_000), ("planar_launch_to_landing_100k", "planar", 100_000)):
        md, blk, wind, desc = make_workload(workload, n, 0)
        eng.set_model(md)
        best = None
        for _ in range(2):
            eng.run_batch(blk, wind, opts=opts)
            c = eng.counters()
            if best is None or c["flight_ms"] < best["flight_ms"]:
                best = c
        ms = best["flight_ms"] + best["rail_ms"]
        sps = best["rk4_steps"] / (best["flight_ms"] * 1e-3)
        out[key] = {"workload": desc, "samples": n, "trajectories_per_s_kernels_only": n / (ms * 1e-3), "rk4_steps_per_s": sps,
                    "mean_rk4_steps_per_trajectory": best["rk4_steps"] / n, "fp64_roofline_frac": sps * FLOP_PER_STEP * 1e-12 / peak_tf}
    return out


def hbm_side(blk, wind, h_out, h_iout, flight_ms, world):
    """The same launch against the HBM roof (to show which roof binds): algorithmic bytes = inputs read once + outputs
    written once, over the flight kernel's time; peak from MEASURED_PEAKS.json (driver-written) or the recipe's fallback."""
    peak, src = 6650.0, "of fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    alg = float(blk.nbytes + wind.nbytes + h_out.nbytes + h_iout.nbytes)
    ach = alg * world / (flight_ms * 1e-3) / 1e9
    return {"algorithmic_bytes_per_launch": alg, "achieved": ach, "peak": peak * world, "unit": "GB/s", "frac": ach / (peak * world),
            "peak_source": src}


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.cpu_samples
    md, blk, wind, desc = make_workload(a.workload, n, 0)
    for _ in range(max(a.warmup, 0)):
        cpu_baseline(md, blk, wind, min(256, n))
    vals, last = [], None
    t_tot = 0.0
    for _ in range(a.steps):
        last, dt = cpu_baseline(md, blk, wind, n)
        vals.append(last["value"]); t_tot += dt
    v = float(n * a.steps / t_tot)
    last["value"] = v
    print(json.dumps({"impl": "reference", "metric": "Monte Carlo trajectories/sec (launch->termination)", "value": v,
                      "unit": "trajectories/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                      "ms_per_step": t_tot / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": desc, "samples_per_step": n, "note": "CPU arm: bounded sample of the same workload"},
                      "cpu_baseline": last,
                      "e2e": {"value": v, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c3", choices=["c3", "planar"])
    ap.add_argument("--samples-per-gpu", type=int, default=100_000)
    ap.add_argument("--cpu-samples", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (larger batch, planar launch->landing set)")
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--refill-threshold", type=int, default=0)
    ap.add_argument("--cold-smem", type=int, default=2)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    if a.impl == "reference":
        return reference_arm(a)

    import torch
    import torch.distributed as dist
    from erpl_monte_carlo_sim_b200 import _abi, _lib, stats as emc_stats

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = a.samples_per_gpu
    md, blk, wind, desc = make_workload(a.workload, n, rank * n)
    eng = _lib.Engine(local)
    eng.set_model(md)
    opts = _lib.run_opts(refill_threshold=a.refill_threshold, block_threads=a.block_threads, blocks_per_sm=a.blocks_per_sm,
                         cold_state_in_smem=a.cold_smem)
    peak_tf, _ = eng.fp64_peak()

    # ---- device-resident inputs/outputs (torch owns the HBM; the engine gets raw device pointers) ----
    d_blk = torch.from_numpy(blk).to(dev); d_wind = torch.from_numpy(wind).to(dev)
    d_out = torch.empty((_abi.OUT_COUNT, n), dtype=torch.float64, device=dev)
    d_iout = torch.empty((_abi.IOUT_COUNT, n), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    last_stats = [None]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        flush.zero_()
        torch.cuda.synchronize()
        eng.run_batch_device(d_blk.data_ptr(), n, d_wind.data_ptr(), wind.shape[1] * 3, d_out.data_ptr(),
                             d_iout.data_ptr(), n, n, opts)
        c = eng.counters()
        # statistics of the batch reduced on the device; for N > 1 the small result blocks are all-reduced over NCCL
        # between the passes (the only collective of the path)
        last_stats[0] = emc_stats.device_statistics(eng, n, out_dev=d_out.data_ptr(), ld=n, distributed=(world > 1))
        return c

    for _ in range(a.warmup):
        step_resident()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    t0 = time.perf_counter()
    flight_ms = rail_ms = 0.0; rk4 = replay = 0
    for _ in range(a.steps):
        c = step_resident()
        flight_ms += c["flight_ms"]; rail_ms += c["rail_ms"]; rk4 += c["rk4_steps"]; replay += c["replay_steps"]
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True; sampler.join(timeout=2)

    # ---- end to end: host pinned buffers through emc_run_batch (H2D + kernels + D2H inside) ----
    p_blk = torch.from_numpy(blk).pin_memory(); p_wind = torch.from_numpy(wind).pin_memory()
    p_out = torch.empty((_abi.OUT_COUNT, n), dtype=torch.float64).pin_memory()
    p_iout = torch.empty((_abi.IOUT_COUNT, n), dtype=torch.int32).pin_memory()
    h_blk, h_wind, h_out, h_iout = p_blk.numpy(), p_wind.numpy(), p_out.numpy(), p_iout.numpy()
    for _ in range(2):
        eng.run_batch(h_blk, h_wind, opts=opts, outputs=(h_out, h_iout))
    barrier()
    t1 = time.perf_counter()
    for _ in range(a.steps):
        flush.zero_()
        eng.run_batch(h_blk, h_wind, opts=opts, outputs=(h_out, h_iout))
    barrier()
    wall_e2e = time.perf_counter() - t1
    parity_hint = bool(np.array_equal(h_iout, d_iout.cpu().numpy()))

    tmax = torch.tensor([wall, wall_e2e, flight_ms, rail_ms], dtype=torch.float64, device=dev)
    tsum = torch.tensor([float(rk4), float(replay)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tsum)
    wall, wall_e2e, flight_ms, rail_ms = tmax.tolist()
    rk4_all, replay_all = tsum.tolist()
    if rank == 0:
        n_total = n * world
        ms_per_step = wall / a.steps * 1e3
        value = n_total / (wall / a.steps)
        e2e_value = n_total / (wall_e2e / a.steps)
        steps_per_s = rk4_all / (flight_ms * 1e-3)                     # over the flight kernel's own device time
        achieved_tf = steps_per_s * FLOP_PER_STEP * 1e-12
        line = {
            "metric": "Monte Carlo trajectories/sec (launch->termination)", "value": value, "unit": "trajectories/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "samples_per_gpu": n, "global_batch": n_total, "seeds": "host-seeded numpy streams, seed = sample index",
                       "l2": "256 MB buffer written between timed iterations", "parallelism": f"sample-sharded x{world}",
                       "launch": {"block_threads": a.block_threads, "blocks_per_sm": a.blocks_per_sm, "refill_threshold": a.refill_threshold, "cold_smem": a.cold_smem}},
            "rk4_steps_per_s": steps_per_s, "mean_rk4_steps_per_trajectory": rk4_all / a.steps / n_total,
            "replayed_steps_per_trajectory": replay_all / a.steps / n_total,
            "kernel_ms_per_step": {"flight": flight_ms / a.steps, "rail": rail_ms / a.steps},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf * world, "unit": "TFLOP/s",
                         "frac": achieved_tf / (peak_tf * world),
                         "traffic": FLIGHT_KERNEL_DRAM_BYTES_100K if (a.workload == "c3" and n == 100_000) else None, "traffic_unit": "bytes/launch (ncu)",
                         "note": "achieved = whole-job RK4 steps/s x 1600 flop (SURVEY 8d) over the flight kernel's CUDA-event time "
                                 "(max over ranks); peak = in-run DFMA-chain microbenchmark (emc_fp64_peak) x n_gpus",
                         "hbm": hbm_side(blk, wind, h_out, h_iout, flight_ms / a.steps, world)},
            "e2e": {"value": e2e_value, "unit": "trajectories/s",
                    "h2d_bytes_per_step": int(blk.nbytes + wind.nbytes), "d2h_bytes_per_step": int(h_out.nbytes + h_iout.nbytes)},
            "gpu_launches": (2 + STATS_LAUNCHES_FUSED) * a.steps * world,
            "clocks": sampler.summary(),
            "statistics": {k: last_stats[0][k] for k in ("n_total", "n_samples", "n_outliers", "apogee_altitude", "range", "flight_time", "landing_ellipse")},
            "e2e_equals_resident": parity_hint,
        }
        if not a.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(md, blk, wind, a.cpu_samples)
        if not a.no_extras and world == 1 and a.workload == "c3":
            line["secondary"] = secondary_measurements(eng, opts, peak_tf)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
