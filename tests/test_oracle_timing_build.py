"""The TIMING build of the oracle (bench.py cpu_baseline / --impl reference: -O3, automatic arrays from a per-thread stack and
NOT initialised, 6-tile halo exchange inside the library) computes what the parity build computes.

Three runs of the same 2-substep dyn_core on the full cube:
  (a) parity build, NumPy halo exchange on the tables of cubed_sphere.py          (the checker of every GPU parity test)
  (b) parity build, fv3o_halo_exchange (C++ gather on the same tables)            -> bit-identical to (a)
  (c) timing build with FV3O_POISON=1: every automatic array starts as NaN         -> no routine reads what it did not write
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import json, sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import harness as H
case = H.Case(12, 6, {fs!r}, state="baroclinic")
oc = H.OracleCube(case, fast={fast}, numpy_halo={numpy_halo})
oc.dyn_core(60.0, 2)
out = {{}}
for t in (1, 4, 6):
    for f in H.regions_state(case.bounds):
        out[f"{{t}}.{{f}}"] = oc.eng[t].get(f).tolist()
json.dump(out, open({dst!r}, "w"))
"""


def _run(tmp_path, tag, fs, fast, numpy_halo, poison):
    dst = str(tmp_path / f"{tag}.json")
    env = dict(os.environ)
    env["FV3O_POISON"] = "1" if poison else "0"
    code = _CHILD.format(root=ROOT, fs=fs, fast=fast, numpy_halo=numpy_halo, dst=dst)
    subprocess.check_call([sys.executable, "-c", code], env=env)
    return {k: np.array(v) for k, v in json.load(open(dst)).items()}


@pytest.mark.parametrize("fs", ["A", "B"])
def test_timing_build_and_library_exchange_match_the_parity_build(built, tmp_path, fs):
    a = _run(tmp_path, "a", fs, False, True, False)
    b = _run(tmp_path, "b", fs, False, False, False)
    c = _run(tmp_path, "c", fs, True, False, True)
    for k in a:
        assert np.array_equal(a[k], b[k]), f"library exchange differs from the NumPy exchange in {k}"
        assert np.all(np.isfinite(c[k])), f"timing build read an uninitialised automatic array ({k})"
        den = max(1.0, float(np.abs(a[k]).max()))
        assert float(np.abs(a[k] - c[k]).max()) / den < 1e-12, k
