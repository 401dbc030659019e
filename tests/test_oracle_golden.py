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


def test_oracle_matches_its_round1_signatures(built):
    """Regression fixture of the oracle itself (tests/golden/make_oracle_selfcheck.py): 2 substeps of the 6-face C12L4 cube,
    flag-sets A and B, hydrostatic and not; sums and 16 sampled values of every prognostic field on two faces."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(os.path.dirname(GOLD), "make_oracle_selfcheck.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    z = np.load(os.path.join(os.path.dirname(GOLD), "oracle_selfcheck.npz"))
    got = {}
    for flagset in ("A", "B"):
        for hydro in (0, 1):
            got.update(mk.run(flagset, hydro))
    assert set(got) == set(z.files)
    for k in z.files:
        ref = z[k]
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(got[k] - ref).max() <= 1e-12 * scale, k
