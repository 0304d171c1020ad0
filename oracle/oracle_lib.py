"""
TEST INFRASTRUCTURE — ctypes binding of oracle/_build/libemc_oracle.so (the plain-C restatement of
the reference hot path, oracle/emc_oracle.c).  Allowed importers: tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from erpl_monte_carlo_sim_b200 import _abi  # noqa: E402  (struct layouts only)

SO = os.path.join(HERE, "_build", "libemc_oracle.so")
_dp = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(HERE, "emc_oracle.c")
    hdr = os.path.join(ROOT, "include", "emc.h")
    if (not force) and os.path.isfile(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return SO
    subprocess.check_call(["make", "-s", "-C", HERE, "CC=gcc"])
    return SO


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(SO)
        L.emc_oracle_derivative.restype = C.c_int
        L.emc_oracle_batch.restype = C.c_int
        L.emc_oracle_tape.restype = C.c_int
        L.emc_oracle_max_threads.restype = C.c_int
        L.orc_interp.restype = C.c_double
        L.orc_interp.argtypes = [C.c_double, _dp, _dp, C.c_int]
        L.orc_gravity.restype = C.c_double
        L.orc_gravity.argtypes = [C.c_void_p, C.c_double]
        L.orc_thrust.restype = C.c_double
        L.orc_thrust.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.orc_atmosphere.argtypes = [C.c_void_p, C.c_double, _dp, _dp, _dp]
        L.orc_mass_properties.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp]
        L.orc_aero_coefficients.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                            C.c_int, C.c_double, _dp]
        L.orc_euler_to_quaternion.argtypes = [C.c_double, C.c_double, C.c_double, _dp]
        L.orc_quaternion_to_euler.argtypes = [_dp, _dp]
        L.orc_rotation_matrix.argtypes = [_dp, _dp]
        _LIB = L
    return _LIB


def _prep(scalars, wind):
    scalars = np.ascontiguousarray(scalars, np.float64)
    if wind is not None and np.size(wind):
        wind = np.ascontiguousarray(wind, np.float64)
    else:
        wind = None
    return scalars, wind


def derivative(md, scalars, wind, t, state, chute):
    L = lib()
    m, keep = _abi.pack_model(md)
    scalars, wind = _prep(scalars, wind)
    n = scalars.shape[1]
    ins = _abi.inputs_struct(scalars, wind, wind_shared=(wind is not None and wind.ndim == 2))
    t = np.ascontiguousarray(t, np.float64)
    state = np.ascontiguousarray(state, np.float64)
    ch = np.ascontiguousarray(chute, np.int32).copy()
    sd = np.empty((n, 14))
    rc = L.emc_oracle_derivative(C.byref(m), C.byref(ins), C.c_int64(n), t.ctypes.data_as(_dp),
                                 state.ctypes.data_as(_dp), ch.ctypes.data_as(C.POINTER(C.c_int32)),
                                 sd.ctypes.data_as(_dp))
    assert rc == 0
    return sd, ch


def batch(md, scalars, wind, n_threads=0, diagnostics=True):
    L = lib()
    m, keep = _abi.pack_model(md)
    scalars, wind = _prep(scalars, wind)
    n = scalars.shape[1]
    ins = _abi.inputs_struct(scalars, wind, wind_shared=(wind is not None and wind.ndim == 2))
    outs, out, iout = _abi.outputs_alloc(n)
    rc = L.emc_oracle_batch(C.byref(m), C.byref(ins), C.c_int64(n), C.byref(outs), C.c_int(n_threads),
                            C.c_int(1 if diagnostics else 0))
    assert rc == 0
    return out, iout


def tape(md, scalars, wind, cap=70000):
    L = lib()
    m, keep = _abi.pack_model(md)
    scalars, wind = _prep(scalars, wind)
    ins = _abi.inputs_struct(scalars, wind, wind_shared=True)
    outs, out, iout = _abi.outputs_alloc(1)
    tp = np.empty((cap, _abi.TAPE_WIDTH))
    ns = C.c_int64(0)
    rc = L.emc_oracle_tape(C.byref(m), C.byref(ins), C.byref(outs), tp.ctypes.data_as(_dp), C.c_int64(cap), C.byref(ns))
    assert rc == 0
    return out, iout, tp[:min(ns.value, cap)]


def series(md, scalars, wind, tape_rows):
    L = lib()
    m, keep = _abi.pack_model(md)
    scalars, wind = _prep(scalars, wind)
    ins = _abi.inputs_struct(scalars, wind, wind_shared=True)
    tape_rows = np.ascontiguousarray(tape_rows, np.float64)
    n = tape_rows.shape[0]
    out = np.empty((_abi.SERIES_COUNT, n))
    L.emc_oracle_series.restype = C.c_int
    rc = L.emc_oracle_series(C.byref(m), C.byref(ins), tape_rows.ctypes.data_as(_dp), C.c_int64(n), out.ctypes.data_as(_dp))
    assert rc == 0
    return out


def max_threads():
    return lib().emc_oracle_max_threads()
