"""CPU: pin the oracle's PPM operators against golden vectors generated from the reference's own
NumPy restatement (docs/examples/tp_core.ipynb -> tests/golden/ppm_notebook.npz, generator
tests/golden/make_ppm_golden.py).  Interior formulas of xppm AND yppm, ord 5, -5, 6, 8, 10."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppm_notebook.npz")


def _cases():
    z = np.load(GOLD)
    keys = sorted(k[:-2] for k in z.files if k.endswith("_q"))
    return z, keys


@pytest.mark.parametrize("ydir", [0, 1])
def test_ppm_matches_reference_notebook(built, ydir):
    lib, _ = H.load_oracle()
    z, keys = _cases()
    assert len(keys) >= 40
    worst = 0.0
    for k in keys:
        q, c, gold = z[k + "_q"], z[k + "_c"], z[k + "_flux"]
        ordv = int(k.split("_ord")[1].split("_")[0].replace("pd", ""))
        iord = -5 if "pd" in k else ordv
        n = q.size
        flux = np.zeros(n + 1)
        dp = C.POINTER(C.c_double)
        rc = lib.fv3o_ppm_periodic(n, q.ctypes.data_as(dp), c.ctypes.data_as(dp), iord, ydir, flux.ctypes.data_as(dp))
        assert rc == 0
        err = np.max(np.abs(flux - gold)) / max(1.0, np.max(np.abs(gold)))
        worst = max(worst, err)
        assert err < 5e-15, (k, err)
    print("worst", worst)
