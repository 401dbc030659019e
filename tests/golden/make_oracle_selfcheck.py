"""Generates tests/golden/oracle_selfcheck.npz: a regression fixture of the ORACLE ITSELF (parity build), so that a later edit of
oracle/ cannot silently change what the CUDA path is compared against.  It does not pin the oracle to the reference (the
reference cannot be run here); it pins the oracle to the state that was validated in round 1 (known-answer test of config 1a,
notebook vectors, invariants, and the GPU parity runs).
   python tests/golden/make_oracle_selfcheck.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402

FIELDS = ("U", "V", "W", "PT", "DELP", "DELZ", "MFX", "CX", "HEAT")


def signature(a):
    a = np.asarray(a, dtype=np.float64).ravel()
    idx = np.linspace(0, a.size - 1, 16).astype(int)
    return np.concatenate([[a.sum(), np.abs(a).sum(), (a * a).sum()], a[idx]])


def run(flagset, hydro):
    case = H.Case(12, 4, flagset, state="baroclinic", flags_override=dict(hydrostatic=hydro))
    oc = H.OracleCube(case, fast=False)
    oc.dyn_core(600.0, 2)
    out = {}
    n = case.n
    for t in (1, 4):
        for f in FIELDS:
            if hydro and f in ("W", "DELZ"):
                continue
            a = oc.eng[t].get(f)
            if a.shape[-1] >= n + 6 and a.shape[-2] >= n + 6:
                a = a[:, 3:-3, 3:-3]          # compute domain only: halo corner blocks are not defined quantities
            out[f"{flagset}_h{hydro}_t{t}_{f}"] = signature(a)
    oc.close()
    return out


def main():
    z = {}
    for flagset in ("A", "B"):
        for hydro in (0, 1):
            z.update(run(flagset, hydro))
    np.savez_compressed(os.path.join(HERE, "oracle_selfcheck.npz"), **z)
    print("wrote", len(z), "signatures")


if __name__ == "__main__":
    main()
