"""MonteCarloAnalyzer — drop-in for the reference's class (monte_carlo.py:17-473).

Same constructor, attributes, `run_monte_carlo` signature and analysis-dict keys.  What changes is
the execution: instead of one `FlightSimulator` per sample in a process pool (monte_carlo.py:67-83)
all dispersed samples go to the CUDA engine as one batch through the C ABI.

Host-seeded mode (this module): the dispersion draws are the reference's own NumPy draws, stream for
stream (monte_carlo.py:156-201,225-335; motor.py:95-125,171-186; environment.py:125-200,218-265), so
the engine sees bit-identical inputs.  tests/test_host_sampling.py pins that against golden inputs
captured from the reference.
"""
from __future__ import annotations

import copy
import os
import time
from collections.abc import Sequence

import numpy as np

from . import _abi, marshal, report, stats
from .motor import LiquidMotor, SolidMotor, is_solid
from .simulator import FlightSimulator, get_engine, summary_extras

PARAM_KEYS = ("initial_position_offset", "initial_velocity_offset", "initial_attitude_offset",
              "initial_angular_velocity_offset", "mass_multiplier", "thrust_multiplier", "wind_speed",
              "wind_direction", "density_multiplier", "random_seed")

# physical bounds of the reference's outlier filter (monte_carlo.py:343-346,383-386)
MAX_REASONABLE_APOGEE = 80000.0
MAX_REASONABLE_RANGE = 200000.0
MAX_REASONABLE_FLIGHT_TIME = 600.0
MIN_REASONABLE_APOGEE = 100.0
THEORETICAL_MAX_ALTITUDE = 1200.0 ** 2 / (2 * 9.81)


class DispersionSet:
    """Struct-of-arrays view of n parameter samples (the reference keeps a list of dicts)."""

    def __init__(self, n):
        self.n = n
        self.pos = np.zeros((n, 3)); self.vel = np.zeros((n, 3)); self.att = np.zeros((n, 3)); self.omega = np.zeros((n, 3))
        self.mass_multiplier = np.ones(n); self.thrust_multiplier = np.ones(n)
        self.wind_speed = np.zeros(n); self.wind_direction = np.zeros(n); self.density_multiplier = np.ones(n)
        self.seed = np.zeros(n, np.int64)

    def as_dict(self, i):
        return {"initial_position_offset": self.pos[i].copy(), "initial_velocity_offset": self.vel[i].copy(),
                "initial_attitude_offset": self.att[i].copy(), "initial_angular_velocity_offset": self.omega[i].copy(),
                "mass_multiplier": float(self.mass_multiplier[i]), "thrust_multiplier": float(self.thrust_multiplier[i]),
                "wind_speed": float(self.wind_speed[i]), "wind_direction": float(self.wind_direction[i]),
                "density_multiplier": float(self.density_multiplier[i]), "random_seed": int(self.seed[i])}

    @classmethod
    def from_dicts(cls, dicts):
        d = cls(len(dicts))
        for i, p in enumerate(dicts):
            d.pos[i] = p["initial_position_offset"]; d.vel[i] = p["initial_velocity_offset"]
            d.att[i] = p["initial_attitude_offset"]; d.omega[i] = p["initial_angular_velocity_offset"]
            d.mass_multiplier[i] = p["mass_multiplier"]; d.thrust_multiplier[i] = p["thrust_multiplier"]
            d.wind_speed[i] = p["wind_speed"]; d.wind_direction[i] = p["wind_direction"]
            d.density_multiplier[i] = p["density_multiplier"]; d.seed[i] = p["random_seed"]
        return d

    def take(self, ids):
        ids = np.asarray(ids, np.int64)
        d = DispersionSet(ids.size)
        for k in ("pos", "vel", "att", "omega", "mass_multiplier", "thrust_multiplier", "wind_speed",
                  "wind_direction", "density_multiplier", "seed"):
            setattr(d, k, getattr(self, k)[ids])
        return d

    def slice(self, lo, hi):
        d = DispersionSet(hi - lo)
        for k in ("pos", "vel", "att", "omega", "mass_multiplier", "thrust_multiplier", "wind_speed",
                  "wind_direction", "density_multiplier", "seed"):
            setattr(d, k, getattr(self, k)[lo:hi])
        return d


class SampleDict(dict):
    """One sample's result dict.  The reference attaches 'trajectory' = {time, altitude, position(n,3)} to every result
    (monte_carlo.py:296-302); here it is served on first access from the downsampled tape the flight kernel recorded
    (BatchRun.trajectory), so building 1e5 result dicts does not move 1e5 time series."""

    def __init__(self, data, run, i):
        super().__init__(data)
        self._run, self._i = run, i

    def __missing__(self, key):
        if key == "trajectory":
            self["trajectory"] = t = self._run.trajectory(self._i)
            return t
        raise KeyError(key)

    def __contains__(self, key):
        return key == "trajectory" or dict.__contains__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default


class LazyAnalysis(dict):
    """The analysis dict of run_monte_carlo.  `parameter_ranges_observed` (monte_carlo.py:425-441) needs the parameter
    dict of every valid sample; it is computed on first access (a 1e7-sample run does not fetch 1e7 x 17 draws unless
    somebody asks).  In a multi-rank job it is computed eagerly, because it takes a collective."""

    def __init__(self, data, lazy):
        super().__init__(data)
        self._lazy = dict(lazy)

    def __missing__(self, key):
        if key in self._lazy:
            self[key] = v = self._lazy.pop(key)()
            return v
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def get(self, key, default=None):
        return self[key] if key in self else default

    def materialize(self):
        for key in list(self._lazy):
            self[key]
        return self

    # anything that enumerates the dict sees every key of the reference's analysis
    def __iter__(self):
        return dict.__iter__(self.materialize())

    def __len__(self):
        return dict.__len__(self.materialize())

    def keys(self):
        return dict.keys(self.materialize())

    def items(self):
        return dict.items(self.materialize())

    def values(self):
        return dict.values(self.materialize())


class SampleResults(Sequence):
    """`analysis['results']` / `analysis['outliers']`: one dict per sample, built on access from the
    engine's SoA summary (the reference materialises every time series of every sample in a list)."""

    def __init__(self, owner, ids, with_reasons=False, length=None):
        """ids: index array, or a zero-argument callable that produces it on first use (it needs the outputs on the host);
        length: the list's length when it is known without them (the device statistics count the valid samples)."""
        self._owner, self._ids_src, self._with_reasons, self._length = owner, ids, with_reasons, length

    @property
    def _ids(self):
        if callable(self._ids_src):
            self._ids_src = self._ids_src()
        if not isinstance(self._ids_src, np.ndarray) or self._ids_src.dtype != np.int64:
            self._ids_src = np.asarray(self._ids_src, np.int64)
        return self._ids_src

    def __len__(self):
        return int(self._length) if self._length is not None else len(self._ids)

    @property
    def sample_indices(self):
        """Indices (into the rank-local batch) of the samples in this list."""
        return self._ids

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        i = int(self._ids[k])
        d = self._owner.sample_result(i)
        if self._with_reasons:
            d["outlier_reasons"] = MonteCarloAnalyzer._outlier_reasons(d["apogee_altitude"], d["range"], d["flight_time"])
        return d


class BatchRun:
    """Everything one Monte Carlo batch produced on this rank: dispersions, inputs and the SoA outputs.  `disp` and
    `scalars` may be given as zero-argument callables: the device-generated modes fetch them only when somebody looks."""

    def __init__(self, analyzer, base_ic, disp, out, iout, altitude_profile, scalars, outputs_resident=False, first_id=0,
                 n=None, engine=None):
        self.analyzer, self.base_ic, self._disp = analyzer, base_ic, disp
        self._out, self._iout, self.altitude_profile, self._scalars = out, iout, altitude_profile, scalars
        self.outputs_resident = outputs_resident      # the engine still holds these outputs in HBM (single chunk)
        self.first_id = int(first_id)                 # global index (= seed) of local sample 0: rank r of a sharded run
        self.n = int(n) if out is None else out.shape[1]
        self._engine = engine                         # out is None: the outputs live in this engine's HBM until somebody looks
        if out is None:
            engine.hold_outputs_for(self)             # ... or until the engine's next batch would overwrite them
        self.tape_ids = np.zeros(0, np.int64)         # local samples whose downsampled trajectory was recorded
        self.tape_rows = np.zeros((0, 2, _abi.BTAPE_WIDTH)); self.tape_count = np.zeros(0, np.int32); self.tape_stride = 0
        self._tape_pos = {}

    def materialize_outputs(self):
        if self._out is None:
            self._out, self._iout = self._engine.fetch_outputs(self.n)

    @property
    def out(self):
        self.materialize_outputs()
        return self._out

    @property
    def iout(self):
        self.materialize_outputs()
        return self._iout

    @property
    def disp(self):
        if callable(self._disp):
            self._disp = self._disp()
        return self._disp

    @property
    def scalars(self):
        if callable(self._scalars):
            self._scalars = self._scalars()
        return self._scalars

    def add_tape(self, ids, rows, count, stride):
        ids = np.asarray(ids, np.int64)
        if self.tape_ids.size and rows.shape[1] != self.tape_rows.shape[1]:
            m = max(rows.shape[1], self.tape_rows.shape[1])
            pad = lambda r: np.concatenate([r, np.full((r.shape[0], m - r.shape[1], r.shape[2]), np.nan)], axis=1)
            rows, self.tape_rows = pad(rows), pad(self.tape_rows)
        base = self.tape_ids.size
        self.tape_ids = np.concatenate([self.tape_ids, ids])
        self.tape_rows = rows if base == 0 else np.concatenate([self.tape_rows, rows])
        self.tape_count = np.concatenate([self.tape_count, np.asarray(count, np.int32)])
        self.tape_stride = stride
        for k, i in enumerate(ids):
            self._tape_pos[int(i)] = base + k

    def ensure_trajectories(self, ids):
        """Record the downsampled trajectories of local samples `ids` that the run did not tape, as ONE extra batch."""
        missing = [int(i) for i in ids if int(i) not in self._tape_pos]
        if missing:
            self.analyzer._tape_samples(self, missing)

    def trajectory(self, i):
        """{'time', 'altitude', 'position' (m,3)} of local sample i from the kernel's downsampled tape (every
        `trajectory_stride`-th stored state and the last one; time since rail exit as in the reference)."""
        if int(i) not in self._tape_pos:
            self.ensure_trajectories([i])
        k = self._tape_pos[int(i)]
        m = int(min(self.tape_count[k], self.tape_rows.shape[1]))
        r = self.tape_rows[k, :m]
        return {"time": r[:, 0].copy(), "altitude": r[:, 3].copy(), "position": r[:, 1:4].copy()}

    def sample_result(self, i):
        O = _abi.OUT
        o = self.out[:, i]
        d = {
            "simulation_id": self.first_id + i, "parameters": self.disp.as_dict(i),
            "apogee_altitude": o[O["apogee_altitude"]], "apogee_time": o[O["apogee_time"]],
            "range": o[O["range"]], "flight_time": o[O["flight_time"]],
            "rail_exit_time": o[O["rail_exit_time"]],
            "rail_exit_position": o[O["rail_exit_x"]:O["rail_exit_z"] + 1].copy(),
            "rail_exit_velocity": o[O["rail_exit_vx"]:O["rail_exit_vz"] + 1].copy(),
            "rail_exit_speed": float(o[O["rail_exit_speed"]]),
            "rail_exit_euler": o[O["rail_exit_roll"]:O["rail_exit_yaw"] + 1].copy(),
            "rail_exit_angle_of_attack": o[O["rail_exit_aoa"]], "rail_exit_sideslip": o[O["rail_exit_sideslip"]],
            "wind_at_exit": o[O["wind_at_exit_u"]:O["wind_at_exit_w"] + 1].copy(),
        }
        d.update(summary_extras(self.out, self.iout, i))
        return SampleDict(d, self, i)

    def full_result(self, i):
        """Re-fly local sample i with the full tape on: the complete `simulate_flight` dict incl. every time series."""
        return self.analyzer.resimulate(self.base_ic, self.disp.as_dict(i))


def shard_range(n, rank, world):
    """Contiguous sample-index range [lo, hi) of rank `rank` among `world` ranks: the GPU-side counterpart of the
    reference's process-pool fan-out (monte_carlo.py:63-83).  Ranges tile [0, n) exactly, sizes differ by at most one."""
    return rank * n // world, (rank + 1) * n // world


class MonteCarloAnalyzer:
    def __init__(self, rocket, motor, atmosphere, wind_model, device=None):
        self.rocket = rocket
        self.motor = motor
        self.atmosphere = atmosphere
        self.wind_model = wind_model
        self.n_cores = os.cpu_count()
        self.base_altitude_profile = None
        self.base_wind_profile = None
        self.uncertainty_params = {
            "initial_position": [0.0, 0.0, 0.0],
            "initial_velocity": [0.1, 0.1, 0.1],
            "initial_attitude": [0.005, 0.005, 0.005],
            "initial_angular_velocity": [0.005, 0.005, 0.005],
            "mass_uncertainty": 0.02,
            "thrust_uncertainty": 0.03,
            "wind_speed_range": [0.0, 5.0],
            "wind_direction_range": [0.0, 2 * np.pi],
            "atmospheric_density_uncertainty": 0.05,
        }
        self.device = device              # None: cuda:LOCAL_RANK inside a torch.distributed job, else cuda:0
        self.chunk_size = 1 << 21         # samples per kernel launch (C4's 1.25 M per GPU is one launch)
        self.run_opts = None
        self.histogram_bins = 0
        # "numpy": the reference's own MT19937 streams drawn on the host (bit-matched inputs)
        # "numpy-device": the same streams regenerated on the GPU (same bits up to the last place of log/sqrt)
        # "philox": counter-based draws on the GPU (same distribution and stream structure)
        # "auto": "numpy" up to host_rng_max samples, "numpy-device" beyond (no per-sample host loop on large runs)
        self.rng = "auto"
        self.host_rng_max = 4096
        self.philox_seed = 0
        # the reference stores the whole trajectory of every sample (monte_carlo.py:296-302); the engine records a
        # downsampled tape {t, x, y, z} of the first `trajectory_samples` samples of a run while they fly (every
        # `trajectory_stride`-th stored state = 0.1 s, plus the last one); any other sample is taped on demand
        self.eager_collectives = True     # torch.distributed jobs: reduce `parameter_ranges_observed` over the ranks inside run_monte_carlo
                                          # (False: on first access, which then has to happen on every rank; it costs a download of the
                                          # shard's outputs and parameters, which a statistics-only campaign does not need)
        self.trajectory_samples = 64
        self.trajectory_stride = 20
        # inside an initialised torch.distributed job the samples are sharded over the ranks (rank r flies the seeds
        # shard_range(n, r, world)) and the statistics are all-reduced over NCCL; False: every rank flies all n samples
        self.shard = True
        self.last_run = None

    def _dev(self):
        if self.device is not None:
            return int(self.device)
        return int(os.environ.get("LOCAL_RANK", "0")) if stats._dist_active() else 0

    def _rank_world(self):
        if self.shard and stats._dist_active():
            import torch.distributed as dist
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _rng_mode(self, n):
        if self.rng == "auto":
            return "numpy" if n <= self.host_rng_max else "numpy-device"
        if self.rng not in ("numpy", "numpy-device", "philox"):
            raise ValueError(f"MonteCarloAnalyzer.rng = {self.rng!r}: expected auto, numpy, numpy-device or philox")
        return self.rng

    # ------------------------------------------------------------------------------------------
    # dispersion draws (host-seeded: the reference's own NumPy streams)
    # ------------------------------------------------------------------------------------------
    def _generate_parameter_samples(self, n_samples):
        """List of dicts, identical to the reference's (monte_carlo.py:156-179)."""
        d = self.draw_parameters(n_samples)
        return [d.as_dict(i) for i in range(n_samples)]

    def _generate_parameter_samples_vectorized(self, n_samples):
        d = self.draw_parameters(n_samples, optimized=True)
        return [d.as_dict(i) for i in range(n_samples)]

    def draw_parameters(self, n_samples, optimized=False, first_seed=0) -> DispersionSet:
        up = self.uncertainty_params
        d = DispersionSet(n_samples)
        s_pos, s_vel = np.asarray(up["initial_position"], float), np.asarray(up["initial_velocity"], float)
        s_att, s_om = np.asarray(up["initial_attitude"], float), np.asarray(up["initial_angular_velocity"], float)
        ws_lo, ws_hi = up["wind_speed_range"]
        wd_lo, wd_hi = up["wind_direction_range"]
        if optimized:
            # one stream, seed 42, drawn sample after sample (monte_carlo.py:181-201)
            rs = np.random.RandomState(42)
            for i in range(n_samples):
                d.pos[i] = rs.normal(0, s_pos); d.vel[i] = rs.normal(0, s_vel)
                d.att[i] = rs.normal(0, s_att); d.omega[i] = rs.normal(0, s_om)
                d.mass_multiplier[i] = rs.normal(1.0, up["mass_uncertainty"])
                d.thrust_multiplier[i] = rs.normal(1.0, up["thrust_uncertainty"])
                d.wind_speed[i] = rs.uniform(ws_lo, ws_hi); d.wind_direction[i] = rs.uniform(wd_lo, wd_hi)
                d.density_multiplier[i] = rs.normal(1.0, up["atmospheric_density_uncertainty"])
                d.seed[i] = i
            return d
        # per-sample stream seeded with the sample index (monte_carlo.py:160-175): 14 normals,
        # 2 uniforms, 1 normal, in that order
        rs = np.random.RandomState(0)
        g = np.empty((n_samples, 15)); u = np.empty((n_samples, 2))
        for i in range(n_samples):
            rs.seed(first_seed + i)
            g[i, :14] = rs.standard_normal(14)
            u[i] = rs.random_sample(2)
            g[i, 14] = rs.standard_normal()
        d.pos[:] = 0 + s_pos * g[:, 0:3]; d.vel[:] = 0 + s_vel * g[:, 3:6]
        d.att[:] = 0 + s_att * g[:, 6:9]; d.omega[:] = 0 + s_om * g[:, 9:12]
        d.mass_multiplier[:] = 1.0 + up["mass_uncertainty"] * g[:, 12]
        d.thrust_multiplier[:] = 1.0 + up["thrust_uncertainty"] * g[:, 13]
        d.wind_speed[:] = ws_lo + (ws_hi - ws_lo) * u[:, 0]
        d.wind_direction[:] = wd_lo + (wd_hi - wd_lo) * u[:, 1]
        d.density_multiplier[:] = 1.0 + up["atmospheric_density_uncertainty"] * g[:, 14]
        d.seed[:] = first_seed + np.arange(n_samples)
        return d

    # ------------------------------------------------------------------------------------------
    # per-sample perturbation -> engine inputs (monte_carlo.py:225-288)
    # ------------------------------------------------------------------------------------------
    def _altitude_grid(self):
        if self.base_wind_profile is not None and self.base_altitude_profile is not None:
            return np.ascontiguousarray(self.base_altitude_profile, np.float64)
        return np.linspace(0, 25000, 100)

    def build_inputs(self, initial_conditions, disp: DispersionSet):
        """(scalars[IN_COUNT][n], wind[n][N][3], altitude grid) for a set of dispersions."""
        n = disp.n
        ic = initial_conditions
        blk = marshal.initial_state_block(
            n, np.asarray(ic.get("position", [0.0, 0.0, 0.0]), float) + disp.pos if "position" in ic else disp.pos,
            np.asarray(ic.get("velocity", [0.0, 0.0, 0.0]), float) + disp.vel if "velocity" in ic else disp.vel,
            np.asarray(ic.get("attitude", [0.0, 0.0, 0.0]), float) + disp.att if "attitude" in ic else disp.att,
            np.asarray(ic.get("angular_velocity", [0.0, 0.0, 0.0]), float) + disp.omega if "angular_velocity" in ic else disp.omega)
        IN = _abi.IN
        dry = self.rocket.dry_mass * disp.mass_multiplier                 # monte_carlo.py:315-316
        prop = self.rocket.propellant_mass * disp.mass_multiplier
        alts = self._altitude_grid()
        n_knots = len(alts)
        # the motor and the wind generator each restart the sample's stream (RandomState(seed)), F11
        rs = np.random.RandomState(0)
        need = max(3 * n_knots, 3)
        gw = np.empty((n, need))
        for i in range(n):
            rs.seed(int(disp.seed[i]))
            gw[i] = rs.standard_normal(need)
        m = self.motor
        if type(m) is LiquidMotor:                                        # motor.py:171-186
            k_t = 1.0 + m.thrust_uncertainty * gw[:, 0]
            k_f = 1.0 + m.mass_flow_uncertainty * gw[:, 1]
            t_vac, t_sl = m.thrust_vacuum * k_t, m.thrust_sea_level * k_t
            mdot = m.mass_flow_rate * k_f
            blk[IN["thrust_a"]] = t_vac
            blk[IN["nozzle_area"]] = (t_vac - t_sl) / 101325.0
            own_burn = m.propellant_mass / mdot
        elif type(m) is SolidMotor:                                       # motor.py:95-125
            k = 1.0 + m.thrust_uncertainty * gw[:, 0]
            mdot = 4.26 * k
            blk[IN["thrust_a"]] = k
            blk[IN["nozzle_area"]] = m.nozzle_exit_area * k
            own_burn = m.burn_time * (1.0 + m.burn_time_uncertainty * gw[:, 1])
        else:                                                             # any duck-typed motor: ask it
            mdot = np.empty(n); own_burn = np.empty(n)
            for i in range(n):
                pm = m.perturb_for_monte_carlo(np.random.RandomState(int(disp.seed[i])))
                mdot[i] = pm.mass_flow_rate; own_burn[i] = pm.burn_time
                blk[IN["nozzle_area"], i] = pm.nozzle_exit_area
                if is_solid(pm):
                    ref = np.asarray(m.thrust_curve_thrust, float)
                    j = int(np.argmax(np.abs(ref)))
                    blk[IN["thrust_a"], i] = np.asarray(pm.thrust_curve_thrust, float)[j] / ref[j]
                else:
                    blk[IN["thrust_a"], i] = pm.thrust_vacuum
        blk[IN["dry_mass"]] = dry
        blk[IN["prop_mass"]] = prop
        blk[IN["mdot"]] = mdot
        with np.errstate(divide="ignore", invalid="ignore"):
            blk[IN["burn_time"]] = np.where(mdot > 0, prop / mdot, own_burn)  # monte_carlo.py:258-260
        blk[IN["cd_scale"]] = 1.0
        g3 = gw[:, :3 * n_knots].reshape(n, n_knots, 3)
        if self.base_wind_profile is not None and self.base_altitude_profile is not None:
            wind = self.wind_model.perturbed_profiles_batch(alts, self.base_wind_profile, g3)   # :271-275
            wind[:, :, 0] += (disp.wind_speed * np.cos(disp.wind_direction))[:, None]            # :277-280
            wind[:, :, 1] += (disp.wind_speed * np.sin(disp.wind_direction))[:, None]
        else:
            wind = self.wind_model.stochastic_profiles_batch(alts, disp.wind_speed, disp.wind_direction, g3)  # :282-288
        return blk, np.ascontiguousarray(wind), alts

    # ------------------------------------------------------------------------------------------
    # device-side dispersions (Philox): same perturbation, draws generated on the GPU
    # ------------------------------------------------------------------------------------------
    def dispersion_struct(self, initial_conditions):
        """(EmcDispersion, keepalive) describing uncertainty_params, the base initial conditions, the nominal rocket and
        motor and the wind generator's per-knot coefficients (environment.py:161-185)."""
        from .environment import _ar1_coefficients
        up, ic, mtr = self.uncertainty_params, initial_conditions, self.motor
        d = _abi.EmcDispersion()
        for name, key in (("base_pos", "position"), ("base_vel", "velocity"), ("base_att", "attitude"), ("base_omega", "angular_velocity")):
            for k, v in enumerate(np.asarray(ic.get(key, [0.0, 0.0, 0.0]), float)):
                getattr(d, name)[k] = v
        for name, key in (("sigma_pos", "initial_position"), ("sigma_vel", "initial_velocity"), ("sigma_att", "initial_attitude"),
                          ("sigma_omega", "initial_angular_velocity")):
            for k, v in enumerate(np.asarray(up[key], float)):
                getattr(d, name)[k] = v
        d.mass_sigma = up["mass_uncertainty"]
        d.wind_speed_lo, d.wind_speed_hi = up["wind_speed_range"]
        d.wind_dir_lo, d.wind_dir_hi = up["wind_direction_range"]
        d.dry_mass, d.propellant_mass = self.rocket.dry_mass, self.rocket.propellant_mass
        d.thrust_vacuum, d.thrust_sea_level, d.mass_flow_rate = mtr.thrust_vacuum, mtr.thrust_sea_level, mtr.mass_flow_rate
        d.nozzle_exit_area, d.motor_propellant_mass, d.motor_burn_time = mtr.nozzle_exit_area, mtr.propellant_mass, mtr.burn_time
        d.thrust_sigma = mtr.thrust_uncertainty
        d.flow_sigma = getattr(mtr, "mass_flow_uncertainty", 0.0)
        d.burn_sigma = getattr(mtr, "burn_time_uncertainty", 0.0)
        d.motor_kind = _abi.MOTOR_SOLID if is_solid(mtr) else _abi.MOTOR_LIQUID
        alts = self._altitude_grid()
        scale, rho, innov = _ar1_coefficients(self.wind_model, alts)
        csv = self.base_wind_profile is not None and self.base_altitude_profile is not None
        d.wind_mode = 1 if csv else 0
        d.n_knots = len(alts)
        shear = np.array([(alts[i] / 10.0) ** self.wind_model.power_law_exponent for i in range(len(alts))])
        base = np.ascontiguousarray(self.base_wind_profile, np.float64) if csv else np.zeros((len(alts), 3))
        rho, innov = np.ascontiguousarray(rho), np.ascontiguousarray(innov)
        dp = _abi._dp
        d.shear, d.base_wind, d.rho, d.innov = (shear.ctypes.data_as(dp), base.ctypes.data_as(dp), rho.ctypes.data_as(dp),
                                                innov.ctypes.data_as(dp))
        return d, (shear, base, rho, innov)

    def philox_parameters(self, n, first_index=0) -> DispersionSet:
        """The parameter samples the device generator draws for indices first_index.. — the draws themselves are fetched
        from the GPU (emc_philox_draws); philox.py holds the host mirror of the same bits (tests/test_philox.py)."""
        idx = np.arange(first_index, first_index + n, dtype=np.uint64)
        g, u = get_engine(self._dev()).philox_draws(self.philox_seed, first_index, n, 15)
        return self._parameters_from_draws(g, u, idx)

    def numpy_device_parameters(self, n, first_seed=0) -> DispersionSet:
        """draw_parameters() from the MT19937 streams the device regenerates (no per-sample host loop)."""
        g, u, dens = get_engine(self._dev()).numpy_draws(first_seed, n, 15)
        g[:, 14] = dens
        return self._parameters_from_draws(g, u, np.arange(first_seed, first_seed + n, dtype=np.uint64))

    def _parameters_from_draws(self, g, u, idx) -> DispersionSet:
        up = self.uncertainty_params
        n = len(idx)
        d = DispersionSet(n)
        d.pos[:] = np.asarray(up["initial_position"], float) * g[:, 0:3]; d.vel[:] = np.asarray(up["initial_velocity"], float) * g[:, 3:6]
        d.att[:] = np.asarray(up["initial_attitude"], float) * g[:, 6:9]; d.omega[:] = np.asarray(up["initial_angular_velocity"], float) * g[:, 9:12]
        d.mass_multiplier[:] = 1.0 + up["mass_uncertainty"] * g[:, 12]; d.thrust_multiplier[:] = 1.0 + up["thrust_uncertainty"] * g[:, 13]
        lo, hi = up["wind_speed_range"]; d.wind_speed[:] = lo + (hi - lo) * u[:, 0]
        lo, hi = up["wind_direction_range"]; d.wind_direction[:] = lo + (hi - lo) * u[:, 1]
        d.density_multiplier[:] = 1.0 + up["atmospheric_density_uncertainty"] * g[:, 14]
        d.seed[:] = idx.astype(np.int64)
        return d

    def run_batch_numpy_device(self, initial_conditions, n, first_seed=0, tape_ids=None) -> BatchRun:
        """Host-seeded semantics without the host: MT19937(seed = sample index) and NumPy's legacy Gaussian regenerated,
        perturbed and flown on the GPU."""
        return self.run_batch_philox(initial_conditions, n, first_index=first_seed, numpy_streams=True, tape_ids=tape_ids)

    def _tape_rows(self):
        """Row capacity of one taped trajectory: every stride-th stored state up to max_time, plus the end points."""
        sim = self._model_simulator()
        dt = min(0.005, float(sim.dt_initial))
        return int(np.ceil(max(float(sim.max_time), 0.0) / dt / max(int(self.trajectory_stride), 1))) + 4

    def _chunk_tape_ids(self, lo, hi, want):
        """Chunk-local indices of the samples in [lo, hi) that are to be taped during the run."""
        if want is not None:
            w = np.asarray(want, np.int64)
            return w[(w >= lo) & (w < hi)] - lo
        return np.arange(lo, min(hi, max(int(self.trajectory_samples), 0)), dtype=np.int64) - lo

    def run_batch_philox(self, initial_conditions, n, first_index=0, numpy_streams=False, tape_ids=None) -> BatchRun:
        """Draw, perturb and fly n samples entirely on the GPU (no host input generation, no input upload)."""
        eng = get_engine(self._dev())
        alts = self._altitude_grid()
        eng.set_model(marshal.model_dict(self.rocket, self.motor, self.atmosphere, self._model_simulator(), alts))
        disp_struct = self.dispersion_struct(initial_conditions)
        single = n <= self.chunk_size
        out = None if single else np.empty((_abi.OUT_COUNT, n))
        iout = None if single else np.empty((_abi.IOUT_COUNT, n), np.int32)
        taped = []
        stride, rows_cap = max(int(self.trajectory_stride), 1), self._tape_rows()
        for lo in range(0, max(n, 1), self.chunk_size):
            hi = min(n, lo + self.chunk_size)
            if numpy_streams:
                eng.generate_inputs_numpy(disp_struct, first_index + lo, hi - lo)
            else:
                eng.generate_inputs(disp_struct, self.philox_seed, first_index + lo, hi - lo)
            ids = self._chunk_tape_ids(lo, hi, tape_ids)
            if ids.size:
                eng.tape_request(ids, stride, rows_cap)
            o, io = eng.run_batch_staged(hi - lo, opts=self.run_opts, download=not single)
            if single:
                out, iout = o, io                    # None: the outputs stay in HBM until somebody looks (BatchRun.out)
            else:
                out[:, lo:hi] = o; iout[:, lo:hi] = io
            if ids.size:
                taped.append((ids + lo,) + eng.tape_fetch())

        def params():
            return self.numpy_device_parameters(n, first_index) if numpy_streams else self.philox_parameters(n, first_index)

        def scalars():
            sc = np.empty((_abi.IN_COUNT, n))
            for lo in range(0, n, self.chunk_size):
                hi = min(n, lo + self.chunk_size)
                if numpy_streams:
                    eng.generate_inputs_numpy(disp_struct, first_index + lo, hi - lo)
                else:
                    eng.generate_inputs(disp_struct, self.philox_seed, first_index + lo, hi - lo)
                sc[:, lo:hi] = eng.staged_inputs(hi - lo, want_wind=False)[0]
            return sc

        run = BatchRun(self, dict(initial_conditions), params, out, iout, alts, scalars, outputs_resident=single, first_id=first_index,
                       n=n, engine=eng)
        run.mode = "numpy-device" if numpy_streams else "philox"
        for ids, rows, cnt in taped:
            run.add_tape(ids, rows, cnt, stride)
        self.last_run = run
        return run

    def _model_simulator(self):
        return FlightSimulator(self.rocket, self.motor, self.atmosphere, self.wind_model, device=self._dev())

    # ------------------------------------------------------------------------------------------
    # running
    # ------------------------------------------------------------------------------------------
    def run_batch(self, initial_conditions, disp: DispersionSet, first_id=0, tape_ids=None) -> BatchRun:
        eng = get_engine(self._dev())
        alts = self._altitude_grid()
        md = marshal.model_dict(self.rocket, self.motor, self.atmosphere, self._model_simulator(), alts)
        eng.set_model(md)
        n = disp.n
        out = np.empty((_abi.OUT_COUNT, n)); iout = np.empty((_abi.IOUT_COUNT, n), np.int32)
        scal = np.empty((_abi.IN_COUNT, n))
        taped = []
        stride, rows_cap = max(int(self.trajectory_stride), 1), self._tape_rows()
        for lo in range(0, n, self.chunk_size):
            hi = min(n, lo + self.chunk_size)
            blk, wind, _ = self.build_inputs(initial_conditions, disp.slice(lo, hi))
            ids = self._chunk_tape_ids(lo, hi, tape_ids)
            if ids.size:
                eng.tape_request(ids, stride, rows_cap)
            o, io = eng.run_batch(blk, wind, opts=self.run_opts)
            out[:, lo:hi] = o; iout[:, lo:hi] = io; scal[:, lo:hi] = blk
            if ids.size:
                taped.append((ids + lo,) + eng.tape_fetch())
        run = BatchRun(self, dict(initial_conditions), disp, out, iout, alts, scal, outputs_resident=(n <= self.chunk_size), first_id=first_id)
        run.mode = "numpy"
        for ids, rows, cnt in taped:
            run.add_tape(ids, rows, cnt, stride)
        self.last_run = run
        return run

    def _tape_samples(self, run: BatchRun, ids):
        """Tape local samples `ids` of a finished run: one extra batch with the tape armed for them.  Host-seeded runs
        re-fly exactly those samples; device-generated runs regenerate and re-fly the index span that covers them."""
        ids = np.unique(np.asarray(ids, np.int64))
        keep_last = self.last_run
        if getattr(run, "mode", "numpy") == "numpy":
            sub = self.run_batch(run.base_ic, run.disp.take(ids), tape_ids=np.arange(ids.size))
            run.add_tape(ids, sub.tape_rows, sub.tape_count, sub.tape_stride)
        else:
            # clusters of ids closer than 4096 apart are regenerated as one span
            cuts = np.flatnonzero(np.diff(ids) > 4096) + 1
            for grp in np.split(ids, cuts):
                lo, hi = int(grp[0]), int(grp[-1]) + 1
                sub = self.run_batch_philox(run.base_ic, hi - lo, first_index=run.first_id + lo,
                                            numpy_streams=(run.mode == "numpy-device"), tape_ids=grp - lo)
                run.add_tape(grp, sub.tape_rows, sub.tape_count, sub.tape_stride)
        run.outputs_resident = False               # the engine's resident block now belongs to the extra batch
        self.last_run = keep_last

    def run_monte_carlo(self, initial_conditions, n_samples=1000, n_processes=None, optimized=False):
        """n_processes is accepted for signature compatibility; the batch runs on the GPU.  Inside an initialised
        torch.distributed job the samples are sharded over the ranks (the fan-out of monte_carlo.py:63-83): every rank
        returns the statistics of the WHOLE job and the result lists of its own shard."""
        if optimized:
            return self.run_optimized_monte_carlo(initial_conditions, n_samples)
        rank, world = self._rank_world()
        lo, hi = shard_range(n_samples, rank, world)
        mode = self._rng_mode(n_samples)
        if mode == "philox":
            run = self.run_batch_philox(initial_conditions, hi - lo, first_index=lo)
        elif mode == "numpy-device":
            run = self.run_batch_numpy_device(initial_conditions, hi - lo, first_seed=lo)
        else:
            run = self.run_batch(initial_conditions, self.draw_parameters(hi - lo, first_seed=lo), first_id=lo)
        return self._analyze_run(run)

    def run_optimized_monte_carlo(self, initial_conditions, n_samples=1000, chunk_size=None):
        t0 = time.time()
        rank, world = self._rank_world()
        lo, hi = shard_range(n_samples, rank, world)
        disp = self.draw_parameters(n_samples, optimized=True)          # one sequential stream (seed 42): draw all, keep the shard
        run = self.run_batch(initial_conditions, disp.slice(lo, hi) if world > 1 else disp, first_id=lo)
        analysis = self._analyze_run(run)
        elapsed = time.time() - t0
        analysis["performance"] = {"total_time": elapsed, "simulations_per_second": n_samples / elapsed,
                                   "cores_used": self.n_cores}
        return analysis

    def run_sweep(self, initial_conditions, pitch_offsets=(0.0,), mass_scales=(1.0,), cd_scales=(1.0,), n_dispersions=1000):
        """Design sweep (BASELINE config C5): every point of the launch-angle x mass x Cd-scale grid flies the same
        `n_dispersions` dispersed samples (common random numbers, seeds 0..n-1) as ONE batch; per point the statistics
        are reduced on the device and the tail is extracted the way the reference's scripts do it by hand: the sample of
        maximum apogee (find_max_apogee.py:7-17) with its outlier diagnostics (analyze_outlier.py:18-25)."""
        eng = get_engine(self._dev())
        alts = self._altitude_grid()
        eng.set_model(marshal.model_dict(self.rocket, self.motor, self.atmosphere, self._model_simulator(), alts))
        disp = self.draw_parameters(n_dispersions)
        grid = [(p, ms, cs) for p in pitch_offsets for ms in mass_scales for cs in cd_scales]
        IN = _abi.IN
        blocks, winds = [], []
        for p, ms, cs in grid:
            ic = dict(initial_conditions)
            att = np.asarray(ic.get("attitude", [0.0, 0.0, 0.0]), float).copy()
            att[1] += p
            ic["attitude"] = att
            blk, wind, _ = self.build_inputs(ic, disp)
            blk[IN["dry_mass"]] *= ms; blk[IN["prop_mass"]] *= ms
            with np.errstate(divide="ignore", invalid="ignore"):
                blk[IN["burn_time"]] = np.where(blk[IN["mdot"]] > 0, blk[IN["prop_mass"]] / blk[IN["mdot"]], blk[IN["burn_time"]])
            blk[IN["cd_scale"]] = cs
            blocks.append(blk); winds.append(wind)
        scal = np.ascontiguousarray(np.concatenate(blocks, axis=1)); wind = np.ascontiguousarray(np.concatenate(winds, axis=0))
        out, iout = eng.run_batch(scal, wind, opts=self.run_opts)
        ptr, ld = eng.resident_outputs()
        O = _abi.OUT
        points = []
        for g, (p, ms, cs) in enumerate(grid):
            lo = g * n_dispersions
            st = stats.device_statistics(eng, n_dispersions, out_dev=ptr + 8 * lo, ld=ld)
            ap = out[O["apogee_altitude"], lo:lo + n_dispersions]
            with np.errstate(invalid="ignore"):
                cand = np.where(ap > 0, ap, 0.0)              # find_max_apogee.py:12-15 starts from max_apogee = 0
            worst = int(np.nanargmax(cand)) if np.any(cand > 0) else -1
            tail = None
            if worst >= 0:
                tail = {"sample": worst, "seed": int(disp.seed[worst]), **summary_extras(out, iout, lo + worst),
                        "apogee_altitude": float(ap[worst]), "flight_time": float(out[O["flight_time"], lo + worst]),
                        "range": float(out[O["range"], lo + worst])}
            points.append({"pitch_offset": p, "mass_scale": ms, "cd_scale": cs, "statistics": st, "max_apogee": tail})
        self.last_run = BatchRun(self, dict(initial_conditions), disp, out, iout, alts, scal, outputs_resident=True)
        return {"grid": grid, "points": points, "n_dispersions": n_dispersions}

    # report writing (monte_carlo.py:475-560): host I/O in the reference's file layout
    def _create_output_directory(self):
        return report.create_output_directory()

    def _save_report(self, analysis, output_dir, max_samples=1000):
        return report.save_report(self, analysis, output_dir, max_samples=max_samples)

    def resimulate(self, initial_conditions, params):
        """The full `simulate_flight` result (time series included) of one dispersed sample."""
        disp = DispersionSet.from_dicts([params])
        blk, wind, alts = self.build_inputs(initial_conditions, disp)
        sim = self._model_simulator()
        rocket, motor = copy.copy(self.rocket), copy.copy(self.motor)     # monte_carlo.py:308-335 deep-copies and mutates
        IN = _abi.IN
        rocket.dry_mass, rocket.propellant_mass = blk[IN["dry_mass"], 0], blk[IN["prop_mass"], 0]
        motor.nozzle_exit_area, motor.mass_flow_rate = blk[IN["nozzle_area"], 0], blk[IN["mdot"], 0]
        motor.burn_time, motor.propellant_mass = blk[IN["burn_time"], 0], blk[IN["prop_mass"], 0]
        if is_solid(self.motor):
            motor.thrust_curve_thrust = np.asarray(self.motor.thrust_curve_thrust, float) * blk[IN["thrust_a"], 0]
        else:
            motor.thrust_vacuum = blk[IN["thrust_a"], 0]
        sim.rocket, sim.motor = rocket, motor
        ic = {"position": blk[0:3, 0].tolist(), "velocity": blk[3:6, 0].tolist(),
              "attitude": (np.asarray(initial_conditions.get("attitude", [0.0, 0.0, 0.0]), float) + disp.att[0]).tolist(),
              "angular_velocity": blk[10:13, 0].tolist()}
        res = sim.simulate_flight(ic, wind[0], alts)
        res["simulation_id"] = int(params.get("random_seed", 0)); res["parameters"] = params
        res["trajectory"] = {"time": res["time"], "altitude": res["altitude"], "position": res["position"].T}
        return res

    # ------------------------------------------------------------------------------------------
    # statistics (monte_carlo.py:337-473)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def outlier_mask(apogee, range_val, flight_time):
        """Vectorised `_filter_physics_outliers` (monte_carlo.py:348-390): True where filtered."""
        with np.errstate(invalid="ignore"):
            bad = ~np.isfinite(apogee) | ~np.isfinite(range_val) | ~np.isfinite(flight_time)
            bad |= (apogee > MAX_REASONABLE_APOGEE) | (apogee < MIN_REASONABLE_APOGEE)
            bad |= range_val > MAX_REASONABLE_RANGE
            bad |= flight_time > MAX_REASONABLE_FLIGHT_TIME
            bad |= apogee > THEORETICAL_MAX_ALTITUDE * 1.2
        return bad

    @staticmethod
    def _outlier_reasons(apogee, range_val, flight_time):
        r = []
        if not np.isfinite(apogee) or not np.isfinite(range_val) or not np.isfinite(flight_time):
            r.append("non-finite values")
        if apogee > MAX_REASONABLE_APOGEE:
            r.append(f"apogee {apogee / 1000:.1f} km > {MAX_REASONABLE_APOGEE / 1000:.1f} km")
        elif apogee < MIN_REASONABLE_APOGEE:
            r.append(f"apogee {apogee:.1f} m < {MIN_REASONABLE_APOGEE:.1f} m")
        if range_val > MAX_REASONABLE_RANGE:
            r.append(f"range {range_val / 1000:.1f} km > {MAX_REASONABLE_RANGE / 1000:.1f} km")
        if flight_time > MAX_REASONABLE_FLIGHT_TIME:
            r.append(f"flight time {flight_time:.1f} s > {MAX_REASONABLE_FLIGHT_TIME:.1f} s")
        if apogee > THEORETICAL_MAX_ALTITUDE * 1.2:
            r.append("apogee exceeds theoretical energy limit")
        return r

    def _filter_physics_outliers(self, results):
        valid, outliers = [], []
        for r in results:
            reasons = self._outlier_reasons(r.get("apogee_altitude", 0), r.get("range", 0), r.get("flight_time", 0))
            if reasons:
                r["outlier_reasons"] = reasons
                outliers.append(r)
            else:
                valid.append(r)
        return valid, outliers

    @staticmethod
    def calc_stats(values):
        values = np.asarray(values, float)
        if values.size == 0:
            nan = float("nan")
            return {"mean": nan, "std": nan, "min": nan, "max": nan, "percentiles": [nan] * 5}
        return {"mean": float(np.mean(values)), "std": float(np.std(values)), "min": float(np.min(values)),
                "max": float(np.max(values)), "percentiles": np.percentile(values, [5, 25, 50, 75, 95]).tolist()}

    def _analyze_results(self, results):
        """List-of-dicts entry kept for callers of the reference's private API (monte_carlo.py:400-473)."""
        initial = [r for r in results if r is not None]
        if len(initial) == 0:
            raise ValueError("No valid simulation results")
        valid, outliers = self._filter_physics_outliers(initial)
        if len(valid) == 0:
            raise ValueError("No physically reasonable simulation results after outlier filtering")
        ap = np.array([r["apogee_altitude"] for r in valid]); rg = np.array([r["range"] for r in valid])
        ft = np.array([r["flight_time"] for r in valid])
        ranges = {}
        for r in valid:
            for key, val in r.get("parameters", {}).items():
                arr = np.array(val)
                if key not in ranges:
                    ranges[key] = {"min": arr.astype(float), "max": arr.astype(float)}
                else:
                    ranges[key]["min"] = np.minimum(ranges[key]["min"], arr)
                    ranges[key]["max"] = np.maximum(ranges[key]["max"], arr)
        for key in ranges:
            ranges[key] = {"min": ranges[key]["min"].tolist(), "max": ranges[key]["max"].tolist()}
        return {"n_samples": len(valid), "n_failed": len(results) - len(initial), "n_outliers": len(outliers),
                "apogee_altitude": self.calc_stats(ap[np.isfinite(ap)]), "range": self.calc_stats(rg[np.isfinite(rg)]),
                "flight_time": self.calc_stats(ft[np.isfinite(ft)]), "results": valid, "outliers": outliers,
                "parameter_ranges_observed": ranges}

    def _analyze_run(self, run: BatchRun):
        """Statistics of a batch: outlier classification, moments, landing ellipse and exact percentiles are reduced
        on the GPU (stats.device_statistics; NCCL all-reduce between the passes in a multi-GPU job, where `run` is this
        rank's shard and the statistics cover every rank's samples)."""
        O = _abi.OUT
        rank, world = self._rank_world()
        if run.n == 0 and world == 1:
            raise ValueError("No valid simulation results")
        eng = get_engine(self._dev())
        if not run.outputs_resident and run.n > 0:
            eng.upload_outputs(run.out)
        st = stats.device_statistics(eng, run.n, histogram_bins=self.histogram_bins, distributed=(world > 1))
        if st["n_total"] == 0:
            raise ValueError("No valid simulation results")
        if st["n_samples"] == 0:
            raise ValueError("No physically reasonable simulation results after outlier filtering")
        split_cache = []

        def split():                                             # per-sample membership for the lazy result lists: needs the
            if not split_cache:                                  # outputs on the host, so it waits until somebody looks
                ap, rg, ft = run.out[O["apogee_altitude"]], run.out[O["range"]], run.out[O["flight_time"]]
                bad = self.outlier_mask(ap, rg, ft)
                split_cache.append((np.flatnonzero(~bad), np.flatnonzero(bad)))
            return split_cache[0]

        def observed_ranges():
            valid_ids = split()[0]
            d = run.disp
            keys, lo, hi = [], [], []
            for key, arr in (("initial_position_offset", d.pos), ("initial_velocity_offset", d.vel),
                             ("initial_attitude_offset", d.att), ("initial_angular_velocity_offset", d.omega),
                             ("mass_multiplier", d.mass_multiplier), ("thrust_multiplier", d.thrust_multiplier),
                             ("wind_speed", d.wind_speed), ("wind_direction", d.wind_direction),
                             ("density_multiplier", d.density_multiplier), ("random_seed", d.seed.astype(float))):
                sel = arr[valid_ids].reshape(valid_ids.size, -1)
                keys.append((key, arr.ndim > 1, sel.shape[1]))
                lo.append(sel.min(axis=0) if valid_ids.size else np.full(sel.shape[1], np.inf))
                hi.append(sel.max(axis=0) if valid_ids.size else np.full(sel.shape[1], -np.inf))
            lo, hi = np.concatenate(lo), np.concatenate(hi)
            if world > 1:
                lo, hi = stats.allreduce_minmax(eng, lo, hi)
            ranges, k = {}, 0
            for key, is_vec, width in keys:
                a, b = lo[k:k + width], hi[k:k + width]
                ranges[key] = {"min": a.tolist(), "max": b.tolist()} if is_vec else {"min": float(a[0]), "max": float(b[0])}
                k += width
            return ranges

        analysis = {"n_samples": st["n_samples"], "n_failed": 0, "n_outliers": st["n_outliers"],
                    "apogee_altitude": st["apogee_altitude"], "range": st["range"], "flight_time": st["flight_time"],
                    "results": SampleResults(run, lambda: split()[0], length=st["n_samples"] if world == 1 else None),
                    "outliers": SampleResults(run, lambda: split()[1], with_reasons=True, length=st["n_outliers"] if world == 1 else None),
                    # engine extras (not in the reference's dict): device-reduced landing ellipse, reasons, histograms
                    "landing_ellipse": st["landing_ellipse"], "outlier_reason_counts": st["outlier_reasons"]}
        if world > 1:
            analysis["shard"] = {"rank": rank, "world_size": world, "first_sample": run.first_id, "n_local": run.n}
        if "histograms" in st:
            analysis["histograms"] = st["histograms"]
        analysis = LazyAnalysis(analysis, {"parameter_ranges_observed": observed_ranges})
        if world > 1 and self.eager_collectives:
            analysis.materialize()                  # the min/max reduction is a collective: every rank takes part now
        return analysis

    # ------------------------------------------------------------------------------------------
    # plots (monte_carlo.py:562-707): host-only matplotlib code on the analysis dict
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _pyplot():
        try:
            import matplotlib
            matplotlib.use("Agg", force=False)
            import matplotlib.pyplot as plt
        except ImportError as e:                                  # the reference imports it at module top (monte_carlo.py:6)
            raise ImportError("MonteCarloAnalyzer.plot_* need matplotlib (the reference lists it in requirements.txt)") from e
        return plt

    @staticmethod
    def _metric_arrays(results):
        """apogee / range / flight time of a result list as arrays (straight from the SoA block when it is ours)."""
        if isinstance(results, SampleResults):
            o, ids, O = results._owner.out, results._ids, _abi.OUT
            return o[O["apogee_altitude"], ids], o[O["range"], ids], o[O["flight_time"], ids]
        return (np.array([r["apogee_altitude"] for r in results], float), np.array([r["range"] for r in results], float),
                np.array([r["flight_time"] for r in results], float))

    def plot_results(self, analysis, save_plots=True):
        """Histograms of apogee, range, flight time and the range-vs-apogee scatter (monte_carlo.py:562-633)."""
        plt = self._pyplot()
        output_dir = None
        _, axes = plt.subplots(2, 2, figsize=(12, 10))
        ap, rg, ft = self._metric_arrays(analysis["results"])
        for ax, vals, label, title in ((axes[0, 0], ap, "Apogee Altitude (m)", "Apogee Altitude Distribution"),
                                       (axes[0, 1], rg, "Range (m)", "Range Distribution"),
                                       (axes[1, 0], ft, "Flight Time (s)", "Flight Time Distribution")):
            ax.hist(vals[np.isfinite(vals)], bins=50, alpha=0.7, edgecolor="black")
            ax.set_xlabel(label); ax.set_ylabel("Frequency"); ax.set_title(title); ax.grid(True, alpha=0.3)
        ok = np.isfinite(ap) & np.isfinite(rg)
        axes[1, 1].scatter(ap[ok], rg[ok], alpha=0.6, s=10)
        axes[1, 1].set_xlabel("Apogee Altitude (m)"); axes[1, 1].set_ylabel("Range (m)")
        axes[1, 1].set_title("Range vs Apogee Altitude"); axes[1, 1].grid(True, alpha=0.3)
        plt.tight_layout()
        if save_plots:
            output_dir = self._create_output_directory()
            plot_path = os.path.join(output_dir, "monte_carlo_distributions.png")
            plt.savefig(plot_path, dpi=300, bbox_inches="tight")
            print(f"Plots saved to: {plot_path}")
            self._save_report(analysis, output_dir)
            print(f"Report saved to: {output_dir}")
        print("\nMonte Carlo Analysis Results:")
        print(f"Number of valid simulations: {analysis['n_samples']}")
        print(f"Number of failed simulations: {analysis['n_failed']}")
        print(f"Number of outlier simulations: {analysis['n_outliers']}")
        for key, title in (("apogee_altitude", "Apogee Altitude"), ("range", "Range")):
            st = analysis[key]
            print(f"\n{title} Statistics:")
            print(f"  Mean: {st['mean']:.1f} m")
            print(f"  Standard Deviation: {st['std']:.1f} m")
            print(f"  95% Confidence Interval: [{st['percentiles'][0]:.1f}, {st['percentiles'][4]:.1f}] m")
        return output_dir

    def _cloud(self, analysis, max_trajectories):
        """The first max_trajectories results with their trajectories; what the run did not tape is taped as ONE batch."""
        results = analysis["results"]
        if isinstance(results, SampleResults):
            results._owner.ensure_trajectories(results._ids[:max_trajectories])
        return results[:max_trajectories]

    def plot_trajectory_cloud(self, analysis, save_plots=True, max_trajectories=50):
        """Altitude-vs-time and ground-track clouds (monte_carlo.py:635-677)."""
        plt = self._pyplot()
        _, (ax1, ax2) = plt.subplots(1, 2, figsize=(15, 6))
        trajectories = self._cloud(analysis, max_trajectories)
        for result in trajectories:
            if "trajectory" in result:
                tr = result["trajectory"]
                ax1.plot(tr["time"], tr["altitude"], alpha=0.3, linewidth=0.5, color="blue")
                if "position" in tr:
                    ax2.plot(tr["position"][:, 0], tr["position"][:, 1], alpha=0.3, linewidth=0.5, color="red")
        ax1.set_xlabel("Time (s)"); ax1.set_ylabel("Altitude (m)")
        ax1.set_title(f"Trajectory Cloud - Altitude vs Time\n({len(trajectories)} trajectories)"); ax1.grid(True, alpha=0.3)
        ax2.set_xlabel("East Position (m)"); ax2.set_ylabel("North Position (m)")
        ax2.set_title(f"Ground Track Cloud\n({len(trajectories)} trajectories)"); ax2.grid(True, alpha=0.3); ax2.axis("equal")
        plt.tight_layout()
        if save_plots:
            output_dir = self._create_output_directory()
            plot_path = os.path.join(output_dir, "monte_carlo_trajectories.png")
            plt.savefig(plot_path, dpi=300, bbox_inches="tight")
            print(f"Trajectory plots saved to: {plot_path}")

    def plot_trajectory_cloud_3d(self, analysis, save_plots=True, max_trajectories=50):
        """3-D trajectory cloud (monte_carlo.py:679-707)."""
        plt = self._pyplot()
        fig = plt.figure(figsize=(10, 8))
        ax = fig.add_subplot(111, projection="3d")
        trajectories = self._cloud(analysis, max_trajectories)
        for result in trajectories:
            if "trajectory" in result and "position" in result["trajectory"]:
                pos = result["trajectory"]["position"]
                ax.plot(pos[:, 0], pos[:, 1], pos[:, 2], alpha=0.3, linewidth=0.5)
        ax.set_xlabel("East Position (m)"); ax.set_ylabel("North Position (m)"); ax.set_zlabel("Altitude (m)")
        ax.set_title(f"3D Trajectory Cloud ({len(trajectories)} trajectories)"); ax.grid(True, alpha=0.3)
        if save_plots:
            output_dir = self._create_output_directory()
            plot_path = os.path.join(output_dir, "monte_carlo_trajectories_3d.png")
            plt.savefig(plot_path, dpi=300, bbox_inches="tight")
            print(f"3D trajectory plot saved to: {plot_path}")
