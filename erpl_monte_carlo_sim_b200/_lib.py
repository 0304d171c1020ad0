"""Loader and thin object wrapper of libemc.so (the C ABI of include/emc.h).

There is no CPU fallback: if the library or a B200-class device is missing every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

from . import _abi

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("EMC_LIB") or os.path.join(PKG_DIR, "libemc.so")    # EMC_LIB: developer override (kernel A/B builds)
CSRC = os.path.join(PKG_DIR, "csrc")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared"]


class EmcError(RuntimeError):
    pass


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/emc_engine.cu for sm_100a into the in-tree libemc.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    srcs = [s for s in srcs if os.path.isfile(s)] + [os.path.join(os.path.dirname(PKG_DIR), "include", "emc.h")]
    if not force and os.path.isfile(SO_PATH) and os.path.getmtime(SO_PATH) >= max(os.path.getmtime(s) for s in srcs):
        return SO_PATH
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH, os.path.join(CSRC, "emc_engine.cu")]
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)          # let nvcc pick the system g++
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise EmcError("nvcc failed building libemc.so:\n" + r.stdout[-4000:])
    return SO_PATH


_LIB = None


def load():
    """dlopen libemc.so and declare the prototypes.  Raises EmcError if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.isfile(SO_PATH):
        raise EmcError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(this engine has no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, i64 = C.c_void_p, C.c_int64
    L.emc_abi_version.restype = C.c_int
    L.emc_last_error.restype = C.c_char_p
    L.emc_last_error.argtypes = [vp]
    L.emc_create.argtypes = [C.POINTER(vp), C.c_int]
    L.emc_destroy.argtypes = [vp]
    L.emc_set_model.argtypes = [vp, C.POINTER(_abi.EmcModel)]
    L.emc_run_batch.argtypes = [vp, C.POINTER(_abi.EmcInputs), i64, C.POINTER(_abi.EmcOutputs), C.POINTER(_abi.EmcRunOpts)]
    L.emc_run_batch_device.argtypes = L.emc_run_batch.argtypes
    L.emc_run_tape.argtypes = [vp, C.POINTER(_abi.EmcInputs), C.POINTER(_abi.EmcOutputs), _dp, i64, C.POINTER(i64)]
    L.emc_derivative_debug.argtypes = [vp, C.POINTER(_abi.EmcInputs), i64, _dp, _dp, _ip, _dp]
    L.emc_get_counters.argtypes = [vp, C.POINTER(_abi.EmcCounters)]
    L.emc_fp64_peak.argtypes = [vp, _dp, _dp]
    L.emc_component_debug.argtypes = [vp, C.c_int, i64, _dp, _dp]
    L.emc_component_debug.restype = C.c_int
    L.emc_fp64_latency.argtypes = [vp, _dp]
    L.emc_fp64_latency.restype = C.c_int
    L.emc_math_debug.argtypes = [vp, C.c_int, i64, _dp, _dp, _dp]
    L.emc_math_debug.restype = C.c_int
    L.emc_scratch.argtypes = [vp, i64, C.POINTER(vp)]
    L.emc_copy_to_host.argtypes = [vp, vp, vp, i64]
    L.emc_copy_to_device.argtypes = [vp, vp, vp, i64]
    L.emc_stats_moments1.argtypes = [vp, vp, i64, i64, vp, vp, vp]
    L.emc_stats_moments2.argtypes = [vp, vp, i64, i64, vp, vp]
    L.emc_stats_select_hist.argtypes = [vp, vp, i64, i64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_int, vp]
    L.emc_stats_select_hist3.argtypes = [vp, vp, i64, i64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int32), vp]
    L.emc_stats_select_hist3.restype = C.c_int
    L.emc_stats_summary.argtypes = [vp, vp, i64, i64, _dp, C.c_int, _dp]
    L.emc_stats_summary.restype = C.c_int
    L.emc_stats_summary_stage.argtypes = [vp, vp, i64, i64, _dp, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), _dp]
    L.emc_stats_summary_stage.restype = C.c_int
    L.emc_stream.argtypes = [vp, C.POINTER(C.c_void_p)]
    L.emc_stream.restype = C.c_int
    L.emc_stats_linear_hist.argtypes = [vp, vp, i64, i64, C.c_int, C.c_double, C.c_double, C.c_int, vp]
    L.emc_generate_inputs.argtypes = [vp, C.POINTER(_abi.EmcDispersion), C.c_uint64, i64, i64, _dp, i64, _dp, vp, i64, vp]
    L.emc_run_batch_staged.argtypes = [vp, i64, C.POINTER(_abi.EmcOutputs), C.POINTER(_abi.EmcRunOpts)]
    L.emc_staged_inputs.argtypes = [vp, i64, _dp, _dp]
    L.emc_philox_draws.argtypes = [vp, C.c_uint64, i64, i64, i64, _dp, _dp]
    L.emc_generate_inputs_numpy.argtypes = [vp, C.POINTER(_abi.EmcDispersion), i64, i64, vp, i64, vp]
    L.emc_numpy_draws.argtypes = [vp, i64, i64, i64, _dp, _dp, _dp]
    for name in ("emc_generate_inputs", "emc_run_batch_staged", "emc_staged_inputs", "emc_philox_draws", "emc_generate_inputs_numpy",
                 "emc_numpy_draws"):
        getattr(L, name).restype = C.c_int
    L.emc_extract_series.argtypes = [vp, C.POINTER(_abi.EmcInputs), _dp, i64, _dp]
    L.emc_extract_series.restype = C.c_int
    L.emc_resident_outputs.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.emc_fetch_outputs.argtypes = [vp, i64, C.POINTER(_abi.EmcOutputs)]
    L.emc_group_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.emc_group_destroy.argtypes = [vp]
    L.emc_group_last_error.restype = C.c_char_p
    L.emc_group_last_error.argtypes = [vp]
    L.emc_group_size.argtypes = [vp]
    L.emc_group_set_model.argtypes = [vp, C.POINTER(_abi.EmcModel)]
    L.emc_group_run_batch.argtypes = [vp, C.POINTER(_abi.EmcInputs), i64, C.POINTER(_abi.EmcOutputs), C.POINTER(_abi.EmcRunOpts)]
    L.emc_group_shard.argtypes = [vp, C.c_int, C.POINTER(i64), C.POINTER(i64)]
    L.emc_group_stats_summary.argtypes = [vp, _dp, C.c_int, _dp]
    L.emc_group_get_counters.argtypes = [vp, C.POINTER(_abi.EmcCounters)]
    L.emc_resident_outputs.restype = C.c_int
    L.emc_upload_outputs.argtypes = [vp, vp, i64, i64]
    L.emc_upload_outputs.restype = C.c_int
    L.emc_tape_request.argtypes = [vp, C.POINTER(C.c_int64), i64, C.c_int32, C.c_int32]
    L.emc_tape_request.restype = C.c_int
    L.emc_tape_fetch.argtypes = [vp, _dp, _ip]
    L.emc_tape_fetch.restype = C.c_int
    L.emc_tape_resident.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(C.c_int32)]
    L.emc_tape_resident.restype = C.c_int
    for name in ("emc_scratch", "emc_copy_to_host", "emc_copy_to_device", "emc_stats_moments1", "emc_stats_moments2",
                 "emc_stats_select_hist", "emc_stats_linear_hist"):
        getattr(L, name).restype = C.c_int
    for name in ("emc_create", "emc_destroy", "emc_set_model", "emc_run_batch", "emc_run_batch_device",
                 "emc_run_tape", "emc_derivative_debug", "emc_get_counters", "emc_fp64_peak"):
        getattr(L, name).restype = C.c_int
    if L.emc_abi_version() != _abi.ABI_VERSION:
        raise EmcError(f"libemc.so ABI {L.emc_abi_version()} != binding ABI {_abi.ABI_VERSION}; rebuild")
    _LIB = L
    return L


def run_opts(refill_threshold=0, block_threads=0, blocks_per_sm=0, nan_fast_forward=True, cold_state_in_smem=2, compaction=False, strict_tail=True, lane_yield=True):
    o = _abi.EmcRunOpts()
    o.refill_threshold = int(refill_threshold)
    o.block_threads = int(block_threads)
    o.blocks_per_sm = int(blocks_per_sm)
    o.nan_fast_forward = 1 if nan_fast_forward else 0
    # EMC_RUN_COMPACTION | EMC_RUN_NO_STRICT_TAIL (ignored by the engine since ABI 2 round 2) | EMC_RUN_NO_YIELD (ABI 3)
    o.flags = (1 if compaction else 0) | (0 if strict_tail else 2) | (0 if lane_yield else 4)
    o.cold_state_in_smem = int(cold_state_in_smem) if cold_state_in_smem else -1    # 2 (default): bookkeeping + base state + RK4 accumulator; 1: bookkeeping; False: registers
    return o


class Engine:
    """One emc_ctx: owns one CUDA device, a stream and its staging buffers."""

    def __init__(self, device: int = 0):
        self._lib = load()
        self._ctx = C.c_void_p()
        rc = self._lib.emc_create(C.byref(self._ctx), int(device))
        if rc != 0:
            msg = self._lib.emc_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise EmcError(f"emc_create failed ({_abi.STATUS.get(rc, rc)}): {msg}")
        self.device = int(device)
        self._model_keep = None
        self.model_dict = None

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.emc_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise EmcError(f"{what} failed ({_abi.STATUS.get(rc, rc)}): {self._lib.emc_last_error(self._ctx).decode()}")

    # -- model ----------------------------------------------------------------------------------
    def set_model(self, md: dict):
        m, keep = _abi.pack_model(md)
        self._check(self._lib.emc_set_model(self._ctx, C.byref(m)), "emc_set_model")
        self._model_keep = (m, keep)
        self.model_dict = md
        self.has_wind = bool(m.has_wind)
        self.n_wind = int(m.n_wind)

    # -- host-buffer path (the drop-in boundary) ------------------------------------------------
    def run_batch(self, scalars, wind=None, opts=None, wind_shared=False, outputs=None):
        """scalars [IN_COUNT][n] float64, wind [n][N][3] (or [N][3] with wind_shared) -> (out, iout).
        `outputs=(out, iout)` lets the caller supply (e.g. pinned) result arrays."""
        self._spill_pending()
        scalars = np.ascontiguousarray(scalars, np.float64)
        n = scalars.shape[1]
        w = None
        if self.has_wind:
            if wind is None:
                raise ValueError("the model has a wind grid: a wind table is required")
            w = np.ascontiguousarray(wind, np.float64)
            if w.ndim == 2:
                wind_shared = True
            if w.shape[-2] != self.n_wind or w.shape[-1] != 3 or (not wind_shared and w.shape[0] != n):
                raise ValueError(f"wind table shape {w.shape} does not match n={n}, n_wind={self.n_wind}")
        ins = _abi.inputs_struct(scalars, w, wind_shared=wind_shared)
        if outputs is None:
            outs, out, iout = _abi.outputs_alloc(n)
        else:
            out, iout = outputs
            assert out.dtype == np.float64 and iout.dtype == np.int32 and out.flags.c_contiguous and iout.flags.c_contiguous
            assert out.shape == (_abi.OUT_COUNT, n) and iout.shape == (_abi.IOUT_COUNT, n)
            outs = _abi.EmcOutputs()
            outs.out, outs.iout, outs.ld = out.ctypes.data, iout.ctypes.data, n
        self._check(self._lib.emc_run_batch(self._ctx, C.byref(ins), n, C.byref(outs),
                                            C.byref(opts) if opts is not None else None), "emc_run_batch")
        return out, iout

    # -- device-buffer path (resident inputs; pointers are raw CUDA device addresses) -----------
    def run_batch_device(self, scalars_ptr, ld, wind_ptr, wind_stride, out_ptr, iout_ptr, out_ld, n, opts=None):
        ins = _abi.EmcInputs()
        ins.scalars = scalars_ptr
        ins.ld = ld
        ins.wind = wind_ptr
        ins.wind_sample_stride = wind_stride
        outs = _abi.EmcOutputs()
        outs.out = out_ptr
        outs.iout = iout_ptr
        outs.ld = out_ld
        self._check(self._lib.emc_run_batch_device(self._ctx, C.byref(ins), n, C.byref(outs),
                                                   C.byref(opts) if opts is not None else None), "emc_run_batch_device")

    def run_tape(self, scalars, wind=None, cap=70000):
        scalars = np.ascontiguousarray(scalars, np.float64).reshape(_abi.IN_COUNT, 1)
        w = np.ascontiguousarray(wind, np.float64) if (self.has_wind and wind is not None) else None
        if self.has_wind and w is None:
            raise ValueError("the model has a wind grid: a wind table is required")
        ins = _abi.inputs_struct(scalars, w, wind_shared=True)
        outs, out, iout = _abi.outputs_alloc(1)
        tape = np.empty((cap, _abi.TAPE_WIDTH), np.float64)
        ns = C.c_int64(0)
        self._check(self._lib.emc_run_tape(self._ctx, C.byref(ins), C.byref(outs), tape.ctypes.data_as(_dp), cap,
                                           C.byref(ns)), "emc_run_tape")
        return out, iout, tape[:ns.value]

    # -- device-side dispersions -----------------------------------------------------------------
    def generate_inputs(self, disp, seed, first_index, n, gauss=None, unif=None, scalars_ptr=None, ld=0, wind_ptr=None):
        """disp = (EmcDispersion, keepalive).  Philox draws unless gauss [n][G] / unif [n][2] are supplied.
        Without device pointers the inputs are staged inside the context (fly them with run_batch_staged)."""
        d, _keep = disp
        g = np.ascontiguousarray(gauss, np.float64) if gauss is not None else None
        u = np.ascontiguousarray(unif, np.float64) if unif is not None else None
        self._check(self._lib.emc_generate_inputs(self._ctx, C.byref(d), C.c_uint64(seed), first_index, n,
                                                  g.ctypes.data_as(_dp) if g is not None else None, g.shape[1] if g is not None else 0,
                                                  u.ctypes.data_as(_dp) if u is not None else None,
                                                  C.c_void_p(scalars_ptr) if scalars_ptr else None, ld,
                                                  C.c_void_p(wind_ptr) if wind_ptr else None), "emc_generate_inputs")
        self._staged_knots = int(d.n_knots)

    def generate_inputs_numpy(self, disp, first_seed, n, scalars_ptr=None, ld=0, wind_ptr=None):
        """The reference's own MT19937 / legacy-Gaussian streams (seed = first_seed + i) regenerated on the device."""
        d, _keep = disp
        self._check(self._lib.emc_generate_inputs_numpy(self._ctx, C.byref(d), first_seed, n,
                                                        C.c_void_p(scalars_ptr) if scalars_ptr else None, ld,
                                                        C.c_void_p(wind_ptr) if wind_ptr else None), "emc_generate_inputs_numpy")
        self._staged_knots = int(d.n_knots)

    def numpy_draws(self, first_seed, n, n_gauss):
        g = np.empty((n, n_gauss), np.float64); u = np.empty((n, 2), np.float64); d = np.empty(n, np.float64)
        self._check(self._lib.emc_numpy_draws(self._ctx, first_seed, n, n_gauss, g.ctypes.data_as(_dp), u.ctypes.data_as(_dp),
                                              d.ctypes.data_as(_dp)), "emc_numpy_draws")
        return g, u, d

    def fetch_outputs(self, n):
        """(out, iout) of the last batch run of n samples, downloaded from the context's resident blocks."""
        outs, out, iout = _abi.outputs_alloc(n)
        self._check(self._lib.emc_fetch_outputs(self._ctx, n, C.byref(outs)), "emc_fetch_outputs")
        return out, iout

    def _spill_pending(self):
        """A batch whose outputs exist only in HBM (BatchRun with lazy outputs) is about to lose them: download first."""
        ref = getattr(self, "_pending", None)
        self._pending = None
        run = ref() if ref is not None else None
        if run is not None:
            run.materialize_outputs()

    def hold_outputs_for(self, run):
        import weakref
        self._pending = weakref.ref(run)

    def run_batch_staged(self, n, opts=None, download=True):
        """Fly the staged samples.  download=False leaves the outputs in HBM only (statistics run there; fetch them later
        with fetch_outputs)."""
        self._spill_pending()
        if download:
            outs, out, iout = _abi.outputs_alloc(n)
        else:
            outs, out, iout = _abi.EmcOutputs(), None, None
            outs.out, outs.iout, outs.ld = None, None, n
        self._check(self._lib.emc_run_batch_staged(self._ctx, n, C.byref(outs), C.byref(opts) if opts is not None else None),
                    "emc_run_batch_staged")
        return out, iout

    # -- downsampled batch tape --------------------------------------------------------------------
    def tape_request(self, samples, stride, max_rows):
        """Arm the next run: record every `stride`-th stored state (+ the last) of the listed samples as {t, x, y, z} rows."""
        idx = np.ascontiguousarray(samples, np.int64)
        self._check(self._lib.emc_tape_request(self._ctx, idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size, int(stride), int(max_rows)),
                    "emc_tape_request")
        self._tape_shape = (idx.size, int(max_rows))

    def tape_fetch(self):
        """(rows[n_sel][max_rows][4], n_rows[n_sel]) of the tape recorded by the last armed run."""
        n_sel, max_rows = self._tape_shape
        rows = np.empty((n_sel, max_rows, _abi.BTAPE_WIDTH), np.float64); cnt = np.empty(n_sel, np.int32)
        self._check(self._lib.emc_tape_fetch(self._ctx, rows.ctypes.data_as(_dp), cnt.ctypes.data_as(_ip)), "emc_tape_fetch")
        return rows, cnt

    def staged_inputs(self, n, want_wind=True):
        sc = np.empty((_abi.IN_COUNT, n), np.float64)
        k = getattr(self, "_staged_knots", 0)
        w = np.empty((n, k, 3), np.float64) if (want_wind and k > 0) else None
        self._check(self._lib.emc_staged_inputs(self._ctx, n, sc.ctypes.data_as(_dp), w.ctypes.data_as(_dp) if w is not None else None),
                    "emc_staged_inputs")
        return sc, w

    def philox_draws(self, seed, first_index, n, n_gauss):
        g = np.empty((n, n_gauss), np.float64); u = np.empty((n, 2), np.float64)
        self._check(self._lib.emc_philox_draws(self._ctx, C.c_uint64(seed), first_index, n, n_gauss, g.ctypes.data_as(_dp),
                                               u.ctypes.data_as(_dp)), "emc_philox_draws")
        return g, u

    def extract_series(self, scalars, wind, tape):
        """Derived per-state series of one flight (simulator.py:496-552) from its tape -> [SERIES_COUNT][n_states]."""
        scalars = np.ascontiguousarray(scalars, np.float64).reshape(_abi.IN_COUNT, 1)
        w = np.ascontiguousarray(wind, np.float64) if (self.has_wind and wind is not None) else None
        ins = _abi.inputs_struct(scalars, w, wind_shared=True)
        tape = np.ascontiguousarray(tape, np.float64)
        n = tape.shape[0]
        series = np.empty((_abi.SERIES_COUNT, n), np.float64)
        self._check(self._lib.emc_extract_series(self._ctx, C.byref(ins), tape.ctypes.data_as(_dp), n, series.ctypes.data_as(_dp)),
                    "emc_extract_series")
        return series

    def derivative_debug(self, scalars, wind, t, state, chute):
        scalars = np.ascontiguousarray(scalars, np.float64)
        n = scalars.shape[1]
        w = np.ascontiguousarray(wind, np.float64) if (self.has_wind and wind is not None and np.size(wind)) else None
        ins = _abi.inputs_struct(scalars, w, wind_shared=(w is not None and w.ndim == 2))
        t = np.ascontiguousarray(t, np.float64)
        state = np.ascontiguousarray(state, np.float64)
        ch = np.ascontiguousarray(chute, np.int32).copy()
        sd = np.empty((n, 14), np.float64)
        self._check(self._lib.emc_derivative_debug(self._ctx, C.byref(ins), n, t.ctypes.data_as(_dp),
                                                   state.ctypes.data_as(_dp), ch.ctypes.data_as(_ip),
                                                   sd.ctypes.data_as(_dp)), "emc_derivative_debug")
        return sd, ch

    COMPONENT_SHAPES = {0: (1, 5), 1: (3, 5), 2: (6, 7), 3: (5, 1)}      # component -> (input rows, output rows)

    def component(self, comp, *columns):
        """Evaluate a model component on the device: columns are broadcast to a common length n -> out[rows][n]."""
        n_in, n_out = self.COMPONENT_SHAPES[comp]
        assert len(columns) == n_in
        cols = np.broadcast_arrays(*[np.atleast_1d(np.asarray(c, np.float64)) for c in columns])
        n = cols[0].size
        blk = np.ascontiguousarray(np.stack([c.ravel() for c in cols], axis=0))
        out = np.empty((n_out, n), np.float64)
        self._check(self._lib.emc_component_debug(self._ctx, comp, n, blk.ctypes.data_as(_dp), out.ctypes.data_as(_dp)),
                    "emc_component_debug")
        return out

    def math_debug(self, op, x, y=None):
        x = np.ascontiguousarray(x, np.float64)
        yy = np.ascontiguousarray(y, np.float64) if y is not None else None
        out = np.empty_like(x)
        self._check(self._lib.emc_math_debug(self._ctx, int(op), x.size, x.ctypes.data_as(_dp),
                                             yy.ctypes.data_as(_dp) if yy is not None else None,
                                             out.ctypes.data_as(_dp)), "emc_math_debug")
        return out

    def resident_outputs(self):
        """(device pointer, leading dimension) of the outputs the last host-buffer run left in HBM."""
        ptr, ld = C.c_void_p(), C.c_int64()
        self._check(self._lib.emc_resident_outputs(self._ctx, C.byref(ptr), C.byref(ld)), "emc_resident_outputs")
        return ptr.value, ld.value

    def upload_outputs(self, out):
        self._spill_pending()
        out = np.ascontiguousarray(out, np.float64)
        assert out.shape[0] == _abi.OUT_COUNT
        self._check(self._lib.emc_upload_outputs(self._ctx, out.ctypes.data, out.shape[1], out.shape[1]), "emc_upload_outputs")

    def counters(self) -> dict:
        c = _abi.EmcCounters()
        self._lib.emc_get_counters(self._ctx, C.byref(c))
        return {k: getattr(c, k) for k, _ in _abi.EmcCounters._fields_}

    def fp64_latency(self):
        c = C.c_double()
        self._check(self._lib.emc_fp64_latency(self._ctx, C.byref(c)), "emc_fp64_latency")
        return c.value

    def fp64_peak(self):
        tf, ms = C.c_double(), C.c_double()
        self._check(self._lib.emc_fp64_peak(self._ctx, C.byref(tf), C.byref(ms)), "emc_fp64_peak")
        return tf.value, ms.value


class EngineGroup:
    """emc_group: several GPUs of one box driven from this ONE process through the C ABI alone (no torch): contiguous sample
    shards flown concurrently, statistics of the whole job reduced with NCCL all-reduces inside libemc.so.  (Jobs launched
    with torchrun use one Engine per rank and torch.distributed instead; see stats.py.)"""

    def __init__(self, devices):
        self._lib = load()
        self._g = C.c_void_p()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = self._lib.emc_group_create(C.byref(self._g), devs, len(devices))
        if rc != 0:
            msg = self._lib.emc_group_last_error(None).decode()
            self._g = C.c_void_p()
            raise EmcError(f"emc_group_create failed ({_abi.STATUS.get(rc, rc)}): {msg}")
        self.devices = [int(d) for d in devices]
        self._model_keep = None

    def close(self):
        if getattr(self, "_g", None) and self._g.value:
            self._lib.emc_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EmcError(f"{what} failed ({_abi.STATUS.get(rc, rc)}): {self._lib.emc_group_last_error(self._g).decode()}")

    def set_model(self, md: dict):
        m, keep = _abi.pack_model(md)
        self._check(self._lib.emc_group_set_model(self._g, C.byref(m)), "emc_group_set_model")
        self._model_keep = (m, keep)
        self.has_wind = bool(m.has_wind)

    def run_batch(self, scalars, wind=None, opts=None):
        """scalars [IN_COUNT][n], wind [n][N][3] -> (out, iout) of all n samples; the shards stay resident per device."""
        scalars = np.ascontiguousarray(scalars, np.float64)
        n = scalars.shape[1]
        w = np.ascontiguousarray(wind, np.float64) if (self.has_wind and wind is not None) else None
        ins = _abi.inputs_struct(scalars, w, wind_shared=(w is not None and w.ndim == 2))
        outs, out, iout = _abi.outputs_alloc(n)
        self._check(self._lib.emc_group_run_batch(self._g, C.byref(ins), n, C.byref(outs), C.byref(opts) if opts is not None else None),
                    "emc_group_run_batch")
        return out, iout

    def shards(self):
        res = []
        for i in range(len(self.devices)):
            a, b = C.c_int64(), C.c_int64()
            self._lib.emc_group_shard(self._g, i, C.byref(a), C.byref(b))
            res.append((a.value, b.value))
        return res

    def stats_summary(self, percentiles=(5, 25, 50, 75, 95)):
        """Raw result block of emc_stats_summary over the whole job: sum[14] | min[3] | max[3] | s2[6] | means[5] | 0 | val[3][2 n_pct]."""
        npct = len(percentiles)
        pct = (C.c_double * npct)(*[float(p) for p in percentiles])
        res = np.empty(32 + 3 * 2 * npct, np.float64)
        self._check(self._lib.emc_group_stats_summary(self._g, pct, npct, res.ctypes.data_as(_dp)), "emc_group_stats_summary")
        return res

    def counters(self) -> dict:
        c = _abi.EmcCounters()
        self._lib.emc_group_get_counters(self._g, C.byref(c))
        return {k: getattr(c, k) for k, _ in _abi.EmcCounters._fields_}
