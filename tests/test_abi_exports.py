"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol include/emc.h
declares, and refuses to run without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import conftest
from erpl_monte_carlo_sim_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "emc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    so = _lib.build_library()
    names = _declared()
    assert {"emc_create", "emc_destroy", "emc_set_model", "emc_run_batch", "emc_run_batch_device", "emc_run_tape",
            "emc_derivative_debug", "emc_last_error"} <= set(names)
    exported = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exp = set(re.findall(r" T (emc_[a-z0-9_]+)", exported))
    missing = [n for n in names if n not in exp]
    assert not missing, f"declared in emc.h but not exported: {missing}"
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n)
    assert lib.emc_abi_version() == _abi.ABI_VERSION


def test_sass_is_sm100a_only():
    so = _lib.build_library()
    out = subprocess.check_output(["cuobjdump", "-lelf", so], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layout_matches_header():
    """sizeof() of the ctypes mirrors against the C compiler's view of include/emc.h."""
    prog = r'''
    #include <stdio.h>
    #include "emc.h"
    int main(){ printf("%zu %zu %zu %zu %zu %d %d %d\n", sizeof(emc_model), sizeof(emc_inputs), sizeof(emc_outputs),
        sizeof(emc_run_opts), sizeof(emc_counters), EMC_IN_COUNT, EMC_OUT_COUNT, EMC_IOUT_COUNT); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe], text=True).split()]
    assert got == [C.sizeof(_abi.EmcModel), C.sizeof(_abi.EmcInputs), C.sizeof(_abi.EmcOutputs), C.sizeof(_abi.EmcRunOpts),
                   C.sizeof(_abi.EmcCounters), _abi.IN_COUNT, _abi.OUT_COUNT, _abi.IOUT_COUNT]


def test_run_option_flags_match_header():
    """The flag bits the Python wrapper sets are the header's EMC_RUN_* values (lane hand-back off = EMC_RUN_NO_YIELD)."""
    from erpl_monte_carlo_sim_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "emc.h")).read()
    bits = {k: int(v) for k, v in re.findall(r"#define (EMC_RUN_\w+) (\d+)", hdr)}
    assert bits == {"EMC_RUN_COMPACTION": 1, "EMC_RUN_NO_STRICT_TAIL": 2, "EMC_RUN_NO_YIELD": 4}
    assert _lib.run_opts().flags == 0
    assert _lib.run_opts(lane_yield=False).flags == bits["EMC_RUN_NO_YIELD"]
    assert _lib.run_opts(compaction=True, lane_yield=False).flags == bits["EMC_RUN_NO_YIELD"] | bits["EMC_RUN_COMPACTION"]
    assert "yielded" in dict(_abi.EmcCounters._fields_) and _abi.ABI_VERSION == int(re.search(r"#define EMC_ABI_VERSION (\d+)", hdr).group(1))


@pytest.mark.skipif(conftest.HAS_CUDA, reason="checks the no-device behaviour")
def test_no_device_is_a_loud_error():
    with pytest.raises(_lib.EmcError, match="EMC_ERR_NO_DEVICE"):
        _lib.Engine(0)
    from erpl_monte_carlo_sim_b200 import FlightSimulator, LiquidMotor, Rocket, StandardAtmosphere, WindModel
    sim = FlightSimulator(Rocket(), LiquidMotor(), StandardAtmosphere(), WindModel())
    with pytest.raises(_lib.EmcError, match="no CPU fallback"):
        sim.simulate_flight({"attitude": [0.0, 0.02, 0.0]})


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "erpl_monte_carlo_sim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in text and "emc_oracle" not in text and "hostseam" not in text.replace("tests/hostseam", ""), f
