"""Post-loop part of dyn_core: del2_cubed (dyn_core.F90:2356-2465; also the omega filter of fv_dynamics.F90:637-642) and the
dissipative-heating update of pt (dyn_core.F90:1300-1356)."""
import numpy as np
import pytest

import harness as H

NG = 3


def _omega_case(n=16, npz=4, flagset="A"):
    case = H.Case(n, npz, flagset, state="baroclinic")
    rng = np.random.default_rng(7)
    fields = []
    for g in case.tiles:
        lon, lat = g.arr["agrid"]
        kk = np.arange(1, npz + 1)[:, None, None]
        f = np.sin(3 * lon + 0.3 * kk)[...] * np.cos(2 * lat)[None] + 0.2 * rng.standard_normal((npz,) + lon.shape)
        fields.append(f)
    return case, fields


def _integral(cube, case, name):
    n = case.n
    return sum(float((cube.eng[t + 1].get(name)[:, NG:NG + n, NG:NG + n] * case.tiles[t].arr["area"][None, NG:NG + n, NG:NG + n]).sum())
               for t in range(6))


@pytest.mark.parametrize("nmax", [1, 2, 3])
def test_oracle_del2_cubed_properties(nmax):
    """A constant stays constant; a noisy field loses variance; the area integral moves only through the corner averaging
    (the filter itself is in flux form)."""
    case, fields = _omega_case()
    n = case.n
    oc = H.OracleCube(case)
    cd = 0.18 * case.tiles[0].da_min
    for t in range(6):
        oc.eng[t + 1].put("OMGA", np.full_like(fields[t], 2.5))
    oc.del2_cubed("OMGA", cd, nmax)
    for t in range(6):
        assert np.abs(oc.eng[t + 1].get("OMGA")[:, NG:NG + n, NG:NG + n] - 2.5).max() < 1e-13
    for t in range(6):
        oc.eng[t + 1].put("OMGA", fields[t])
    i0 = _integral(oc, case, "OMGA")
    v0 = sum(float(np.var(fields[t][:, NG:NG + n, NG:NG + n])) for t in range(6))
    a0 = sum(float(np.abs(fields[t][:, NG:NG + n, NG:NG + n] * case.tiles[t].arr["area"][None, NG:NG + n, NG:NG + n]).sum()) for t in range(6))
    oc.del2_cubed("OMGA", cd, nmax)
    v1 = sum(float(np.var(oc.eng[t + 1].get("OMGA")[:, NG:NG + n, NG:NG + n])) for t in range(6))
    i1 = _integral(oc, case, "OMGA")
    oc.close()
    assert v1 < 0.9 * v0
    assert abs(i1 - i0) < 2e-3 * a0      # 8 corners x 3 averaged cells of 6 x 16^2 cells


def test_oracle_dcon_heating_warms_and_is_bounded():
    """Flag-set B (d_con = 1): after dyn_core the accumulated heat source has been turned into a temperature tendency
    (heat_source holds K per step, |.| limited by delt_max * bdt scaled by the sponge factors) and pt changed accordingly."""
    case = H.Case(12, 5, "B", state="baroclinic")
    a = H.OracleCube(case)
    bdt = 600.0
    a.dyn_core(bdt, 2)
    # same run without the post-loop step
    b = H.OracleCube(case)
    b.dcon_heating = lambda bdt: None
    b.dyn_core(bdt, 2)
    n = case.n
    changed = 0.0
    for t in a.tiles:
        pa, pb = a.eng[t].get("PT")[:, NG:NG + n, NG:NG + n], b.eng[t].get("PT")[:, NG:NG + n, NG:NG + n]
        pkz = a.eng[t].get("PKZ")
        d = (pa - pb) * pkz                    # K
        lim = abs(bdt) * case.flags["delt_max"] * np.array([0.1, 0.5, 1, 1, 1])[:, None, None]
        assert (np.abs(d) <= lim * (1 + 1e-12)).all()
        changed = max(changed, float(np.abs(d).max()))
        assert np.isfinite(pa).all()
    a.close(); b.close()
    assert changed > 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("nmax", [1, 2, 3])
def test_cuda_del2_cubed_matches_oracle(nmax):
    case, fields = _omega_case(n=32, npz=3)
    oc, gc = H.OracleCube(case), H.CudaCube(case)
    cd = 0.18 * case.tiles[0].da_min
    for t in range(6):
        oc.eng[t + 1].put("OMGA", fields[t]); gc.eng[t + 1].put("OMGA", fields[t])
    oc.del2_cubed("OMGA", cd, nmax); gc.del2_cubed("OMGA", cd, nmax)
    b = case.bounds
    for t in oc.tiles:
        err = H.compare(oc.eng[t], gc.eng[t], {"OMGA": (b["is_"], b["ie"], b["js"], b["je"])})["OMGA"]
        assert err < 1e-13, (t, err)
    oc.close(); gc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("hydro", [0, 1])
def test_cuda_dyn_core_with_dcon_heating(hydro):
    """Flag-set B through fv3_dyn_core incl. the post-loop heating; pt, heat_source and pkz against the oracle."""
    case = H.Case(16, 6, "B", state="baroclinic", flags_override=dict(hydrostatic=hydro))
    oc, gc = H.OracleCube(case), H.CudaCube(case)
    oc.dyn_core(800.0, 2); gc.dyn_core(800.0, 2)
    b = case.bounds
    reg = (b["is_"], b["ie"], b["js"], b["je"])
    for t in oc.tiles:
        res = H.compare(oc.eng[t], gc.eng[t], {"PT": reg, "HEAT": reg, "PKZ": reg, "DELP": reg})
        for f, e in res.items():
            assert e < 1e-10, (t, f, e)   # multi-substep tolerance of tests/test_gpu_parity.py (measured: heat_source 1.4e-12)
    hs = gc.eng[1].get("HEAT")
    assert np.abs(hs).max() > 0.0
    oc.close(); gc.close()


def _theta_case(hydro=0, use_cond=0):
    case = H.Case(12, 5, "A", state="baroclinic", flags_override=dict(hydrostatic=hydro, use_cond=use_cond))
    rng = np.random.default_rng(3)
    qv = [1e-3 * (1.0 + rng.random((5,) + g.arr["agrid"][0].shape)) for g in case.tiles]
    qc = [1e-4 * rng.random((5,) + g.arr["agrid"][0].shape) for g in case.tiles]
    return case, qv, qc


def _load_theta(e, case, t, qv, qc, T):
    e.put("WORK_Q", qv[t]); e.put("QCON", qc[t]); e.put("PT", T)
    if case.flags["hydrostatic"]:
        n = case.n
        e.put("PKZ", np.full((5, n, n), 1.0) * (np.linspace(5.0, 45.0, 5)[:, None, None]))


@pytest.mark.parametrize("hydro,use_cond", [(0, 0), (0, 1), (1, 0)])
def test_oracle_pt_to_theta_matches_the_formulas(hydro, use_cond):
    """fv_dynamics.F90:303-328, 377-398 against the same expressions in NumPy."""
    case, qv, qc = _theta_case(hydro, use_cond)
    lib = H.load_oracle()
    zvir = 0.6078
    cst = case.consts
    for t in (0, 3):
        e = case.engine(lib, t + 1)
        case.load_state(e, t + 1)
        T = 250.0 + 30.0 * np.cos(case.tiles[t].arr["agrid"][1])[None] + np.arange(5)[:, None, None]
        _load_theta(e, case, t, qv, qc, T)
        pkz_in = e.get("PKZ")
        e.call("pt_to_theta", zvir)
        n = case.n
        sl = (slice(None), slice(NG, NG + n), slice(NG, NG + n))
        d1 = zvir * qv[t][sl]
        delp, delz = e.get("DELP")[sl], e.get("DELZ")
        if hydro:
            pkz = pkz_in
        else:
            pkz = np.exp(cst["kappa"] * np.log(-cst["rdgas"] / cst["grav"] * delp * T[sl] * (1 + d1) / delz))
        want = T[sl] * (1 + d1) * ((1 - qc[t][sl]) if use_cond else 1.0) / pkz
        assert np.abs(e.get("PT")[sl] / want - 1).max() < 1e-14
        assert np.abs(e.get("PKZ") / pkz - 1).max() < 1e-14
        assert np.abs(e.get("DP1")[sl] - d1).max() == 0.0
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("hydro,use_cond", [(0, 0), (0, 1), (1, 0)])
def test_cuda_pt_to_theta_matches_oracle(hydro, use_cond):
    case, qv, qc = _theta_case(hydro, use_cond)
    from gfdl_atmos_cubed_sphere_b200 import abi
    eo, eg = case.engine(H.load_oracle(), 2), case.engine(abi.load_library(), 2)
    T = 250.0 + 30.0 * np.cos(case.tiles[1].arr["agrid"][1])[None] + np.arange(5)[:, None, None]
    for e in (eo, eg):
        case.load_state(e, 2)
        _load_theta(e, case, 1, qv, qc, T)
        e.call("pt_to_theta", 0.6078)
    b = case.bounds
    reg = (b["is_"], b["ie"], b["js"], b["je"])
    res = H.compare(eo, eg, {"PT": reg, "PKZ": reg, "DP1": reg})
    for f, err in res.items():
        assert err < 1e-13, (f, err)
    eo.close(); eg.close()
