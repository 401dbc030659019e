"""CPU: the C-ABI library loads without a GPU and exports every symbol include/fv3_dyncore.h declares;
the ctypes mirrors have the sizes the library was compiled with; the product fails loudly without CUDA."""
import ctypes as C
import os
import re

import pytest

import harness as H
from gfdl_atmos_cubed_sphere_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib(built):
    if not os.path.exists(built.LIB):
        pytest.skip("CUDA library not built (no nvcc here)")
    return C.CDLL(built.LIB)


def test_header_symbols_are_exported(built):
    lib = _lib(built)
    hdr = open(os.path.join(ROOT, "include", "fv3_dyncore.h")).read()
    names = set(re.findall(r"\b(fv3_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 30
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_sizes_match(built):
    lib = _lib(built)
    assert lib.fv3_abi_sizeof(0) == C.sizeof(abi.Bounds)
    assert lib.fv3_abi_sizeof(1) == C.sizeof(abi.Grid)
    assert lib.fv3_abi_sizeof(2) == C.sizeof(abi.Flags)
    assert lib.fv3_abi_sizeof(3) == C.sizeof(abi.State)


def test_no_cpu_fallback(built):
    """Without a CUDA device fv3_create must fail (non-zero), never silently compute on the CPU."""
    _lib(built)
    if H.have_gpu():
        pytest.skip("GPU present")
    case = H.Case(8, 2, "A")
    with pytest.raises(RuntimeError, match="create failed"):
        case.engine(abi.load_library(), 1)


def test_oracle_is_not_reachable_from_the_package():
    pkg = os.path.join(ROOT, "gfdl_atmos_cubed_sphere_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "libfv3_oracle" not in src and "fv3o_" not in src, f


def test_unbuilt_sw_test_cases_are_rejected(built):
    """Only test_case 1 of the SW_DYNAMICS build exists (BASELINE config 1a); other values are an argument error (-2), with or
    without a GPU -- the check precedes any device work."""
    _lib(built)
    case = H.Case(8, 1, "A", flags_override=dict(sw_test_case=2))
    with pytest.raises(RuntimeError, match="rc=-2"):
        case.engine(abi.load_library(), 1)
