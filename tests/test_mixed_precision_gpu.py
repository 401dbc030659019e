"""-m gpu: BASELINE config 5 ("fp32 transport / fp64 Riemann"), fv3_set_transport_fp32.

The PPM sweeps of d_sw's interior tiles compute in fp32 from the fp64 fields; the fluxes are applied to the fp64 prognostics in
fp64.  What is checked:
  * the mode is really engaged (results differ from the fp64 run) and stays within single-precision distance of it and of the
    oracle -- the bound is the fp32 rounding of a FLUX (6e-8) times the fraction of a cell's content that crosses a face in one
    substep (a Courant number <~ 0.1..0.3), i.e. ~1e-8 per substep on delp, pt; it is written below with a margin;
  * the update is still conservative: both cells of a face apply the same rounded flux (sw_core.F90:1059-1060), also when the
    face is shared by two tiles (every tile runs the same instantiation in this mode).  The one place where two parties compute a
    flux separately is a CUBE EDGE: the two faces evaluate the same expressions mirrored, which agree to the round-off of the
    arithmetic type -- 1e-16 in fp64, 6e-8 of the edge fluxes here (as in the reference's own 32-bit build).  Global mass
    therefore drifts by ~1e-13 per substep (measured 4e-13 after four) instead of 1e-15; the bound below is 2e-12, four orders of
    magnitude under what inconsistent tile-boundary fluxes would give;
  * nothing but the sweeps changed: with the mode switched off again the run is bit-identical to a run that never had it on.
C96 is the smallest face with interior tiles (4 x 4 tiles of 26 cells: the 2 x 2 in the middle).
"""
import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu

N, NPZ = 96, 6
FIELDS = ("DELP", "PT", "W", "U", "V")
# fp32 sweeps vs fp64 sweeps after two substeps (max |a - b| / max |b|): delp, pt carry the flux rounding diluted by the Courant
# number (measured 2.4e-8 after one substep of 225 s at C96, 5e-8 after two: profiles/diag_fp32.py).
TOL32 = {"DELP": 2e-7, "PT": 2e-7}
# The winds are judged in ABSOLUTE terms.  A relative pressure perturbation e of the fp32 fluxes becomes an acoustic response of
# dt * (p e / rho) / dx = 225 s * 3e-8 * 8e4 m2/s2 / 1e5 m ~ 5e-6 m/s in u, v and dt * g * e ~ 3e-6 m/s in w (measured 9e-6, 2e-6);
# on top of that a monotonicity constraint flips at isolated points (the limiters select with min / max / sign), which moves u, v
# by ~1e-4 m/s there (measured 1.7e-4 after two substeps, not growing with the third).  Bounds with a margin of ~5:
TOL32_ABS = {"U": 1e-3, "V": 1e-3, "W": 5e-5}


def _box(name):
    return (1, N + (name == "V"), 1, N + (name == "U"))


def _run(case, fp32, substeps=2):
    gc = H.CudaCube(case)
    if fp32 is not None:
        gc.set_transport_fp32(fp32)
    gc.dyn_core(225.0 * substeps, substeps)
    out = {t: {f: H.sub(gc.eng[t], f, gc.eng[t].get(f), *_box(f)) for f in FIELDS} for t in gc.tiles}
    return gc, out


def _err(a, b):
    return float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-300)


def _worst(res, ref_of):
    rel = {f: 0.0 for f in TOL32}
    ab = {f: 0.0 for f in TOL32_ABS}
    for t in res:
        for f in FIELDS:
            ref = ref_of(t, f)
            if f in rel:
                rel[f] = max(rel[f], _err(res[t][f], ref))
            else:
                ab[f] = max(ab[f], float(np.abs(res[t][f] - ref).max()))
    return rel, ab


def _check(rel, ab):
    for f in TOL32:
        assert rel[f] < TOL32[f], (f, rel[f])
    for f in TOL32_ABS:
        assert ab[f] < TOL32_ABS[f], (f, ab[f])


@pytest.mark.parametrize("flagset", ["A", "B"])
def test_fp32_transport_is_engaged_and_close_to_fp64(flagset):
    case = H.Case(N, NPZ, flagset, state="baroclinic")
    g64, r64 = _run(case, None)
    g32, r32 = _run(case, True)
    rel, ab = _worst(r32, lambda t, f: r64[t][f])
    print("fp32 sweeps vs fp64 sweeps: relative", {k: f"{v:.2e}" for k, v in rel.items()}, "absolute (m/s)", {k: f"{v:.2e}" for k, v in ab.items()})
    assert rel["DELP"] > 1e-13, "fp32 mode produced bit-identical delp: the fp32 kernels did not run"
    _check(rel, ab)
    g64.close(); g32.close()


def test_fp32_transport_conserves_mass_to_fp64_roundoff():
    case = H.Case(N, NPZ, "A", state="baroclinic")
    gc = H.CudaCube(case)
    gc.set_transport_fp32(True)

    def mass():
        tot = 0.0
        for t in gc.tiles:
            d = H.sub(gc.eng[t], "DELP", gc.eng[t].get("DELP"), 1, N, 1, N)
            tot += float(np.sum(d * case.tiles[t - 1].arr["area"][None, 3:-3, 3:-3]))
        return tot
    m0 = mass()
    gc.dyn_core(900.0, 4)
    m1 = mass()
    print(f"relative mass drift after 4 substeps with fp32 sweeps: {abs(m1 - m0) / m0:.2e}")
    assert abs(m1 - m0) / m0 < 2e-12, (m0, m1)
    gc.close()


def test_fp32_transport_against_the_oracle_and_switch_off():
    """one substep against the fp64 oracle at single-precision tolerance; switching the mode off restores the fp64 path bit for bit"""
    case = H.Case(N, 3, "A", state="baroclinic")
    oc = H.OracleCube(case)
    oc.dyn_core(225.0, 1)
    g32, r32 = _run(case, True, substeps=1)
    rel, ab = _worst(r32, lambda t, f: H.sub(oc.eng[t], f, oc.eng[t].get(f), *_box(f)))
    print("fp32 sweeps vs oracle: relative", {k: f"{v:.2e}" for k, v in rel.items()}, "absolute (m/s)", {k: f"{v:.2e}" for k, v in ab.items()})
    _check(rel, ab)
    g32.close()
    g_off, r_off = _run(case, False, substeps=1)
    g_def, r_def = _run(case, None, substeps=1)
    for t in g_off.tiles:
        for f in FIELDS:
            assert np.array_equal(r_off[t][f], r_def[t][f])
    g_off.close(); g_def.close(); oc.close()


@pytest.mark.parametrize("fp32", [0, 1])
def test_single_face_mass_budget_closes_across_tile_boundaries(fp32):
    """sum over a face of area * (delp_new - delp_old) == the mass fluxes through the face's boundary, to fp64 round-off.  Every
    interior face of the 4 x 4 tiling must cancel for that -- in particular the faces SHARED by an interior tile and a frame tile,
    which two different kernel instantiations evaluate: bit-identical in fp32 mode by construction (strict float, tp_line.cuh);
    a float-rounding mismatch there would leave a residual of ~1e-11 of the face's mass (3 x 96 x 2 x npz faces x 6e-8 of a flux)."""
    from gfdl_atmos_cubed_sphere_b200 import abi
    case = H.Case(N, NPZ, "A", state="baroclinic")
    e = case.engine(abi.load_library(), 1)
    case.load_state(e, 1)
    e.lib.fv3_set_transport_fp32(e.ctx, fp32)
    dt = 225.0
    e.call("c_sw", 0.5 * dt)
    area = case.tiles[0].arr["area"][3:-3, 3:-3]
    d0 = H.sub(e, "DELP", e.get("DELP"), 1, N, 1, N).copy()
    e.put("MFX", np.zeros(e.shape("MFX")))
    e.put("MFY", np.zeros(e.shape("MFY")))
    e.call("d_sw", dt)
    e.sync()
    d1 = H.sub(e, "DELP", e.get("DELP"), 1, N, 1, N)
    mfx, mfy = e.get("MFX"), e.get("MFY")
    fx = lambda i: H.sub(e, "MFX", mfx, i, i, 1, N).astype(np.longdouble).sum()
    fy = lambda j: H.sub(e, "MFY", mfy, 1, N, j, j).astype(np.longdouble).sum()
    lhs = ((d1.astype(np.longdouble) - d0) * area[None]).sum()
    rhs = fx(1) - fx(N + 1) + fy(1) - fy(N + 1)
    mass = float((d0 * area[None]).sum())
    print(f"fp32={fp32}: budget residual {float(lhs - rhs):.3e} of face mass {mass:.3e} = {abs(float(lhs - rhs)) / mass:.2e}")
    assert abs(float(lhs - rhs)) / mass < 1e-14
    e.close()
