"""-m gpu: the vertical remap (csrc/remap.cu) against the oracle (oracle/remap.cpp) through the C ABI.

Tolerances: 1e-12 for one call of a column operator / of Lagrangian_to_Eulerian on identical inputs (same arithmetic up to FMA
contraction; the branchy limiters are exact), 1e-9 for the dyn_core -> remap -> dyn_core -> remap sequence of a k_split loop
(w: 1e-5, see the test).

Schemes 11 and 12 switch on |2 a1 - (a2 + a3)| > |a2 - a3| (fv_operators.F90:706-714), which is an exact TIE wherever an edge value
was clipped to the layer mean; the outcome then hangs on the last bit of the inputs and moves the result by 1e-4.  remap.cu is
therefore compiled without FMA contraction (bit-identical to the oracle on identical inputs: the column-operator test below
covers 11 and 12 in every mode), and Lagrangian_to_Eulerian is compared with kord_tm in 8, 9, 10, 13 -- the temperature path goes
through exp / log, where the device and libm differ in the last bit -- while the winds, w and the tracer also run 11 and 12."""
import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu
TOL = 1e-12
N, NPZ = 24, 16
KORDS = [8, 9, 10, 11, 12, 13, 14, 15]
STATE = ("PT", "DELP", "DELZ", "W", "U", "V", "PE", "PELN", "PK", "PKZ", "OMGA", "WORK_Q", "WS")


def _regions(b, hydro=False):
    is_, ie, js, je = b["is_"], b["ie"], b["js"], b["je"]
    r = {"PT": (is_, ie, js, je), "DELP": (is_, ie, js, je), "U": (is_, ie, js, je + 1), "V": (is_, ie + 1, js, je),
         "PE": (is_, ie, js, je), "PELN": (is_, ie, js, je), "PK": (is_, ie, js, je), "PKZ": (is_, ie, js, je),
         "OMGA": (is_, ie, js, je), "WORK_Q": (is_, ie, js, je)}
    if not hydro:
        r.update({"W": (is_, ie, js, je), "DELZ": (is_, ie, js, je)})
    return r


def _pair(substeps=2, **over):
    """oracle and CUDA cubes after `substeps` acoustic substeps; the CUDA state is then overwritten with the oracle's, so a
    remap call is compared on identical inputs"""
    case = H.Case(N, NPZ, "A", state="baroclinic", flags_override=over or None)
    oc, gc = H.OracleCube(case), H.CudaCube(case)
    oc.dyn_core(450.0 * substeps, substeps)
    gc.dyn_core(450.0 * substeps, substeps)
    rng = np.random.default_rng(11)
    for t in oc.tiles:
        q = oc.eng[t].get("WORK_Q")
        q[...] = np.abs(oc.eng[t].get("PT")) * rng.uniform(0.0, 1.0, q.shape) ** 4      # a positive tracer with sharp features
        oc.eng[t].put("WORK_Q", q)
        om = oc.eng[t].get("OMGA"); om[...] = rng.normal(0.0, 0.2, om.shape); oc.eng[t].put("OMGA", om)
        for f in STATE:
            gc.eng[t].put(f, oc.eng[t].get(f))
    return case, oc, gc


def _assert(res, tol):
    bad = {k: v for k, v in res.items() if not (v <= tol)}
    assert not bad, f"parity exceeded {tol}: {bad}"


@pytest.mark.parametrize("kord", KORDS)
def test_column_operators_match_the_oracle(kord):
    case, oc, gc = _pair()
    b = case.bounds
    reg = {"WORK_Q": (b["is_"], b["ie"], b["js"], b["je"])}
    for mode, iv in [(0, 1), (1, 1), (1, -1), (1, -2), (2, 0), (0, 0), (1, 2)]:
        for t in (1, 3, 6):
            eo, eg = oc.eng[t], gc.eng[t]
            q0 = eo.get("WORK_Q")
            for e in (eo, eg):
                e.call("remap_work_q", mode, iv, kord, 1.0 if mode == 0 else 0.0)
            _assert(H.compare(eo, eg, reg), TOL)
            for e in (eo, eg):
                e.put("WORK_Q", q0)
    oc.close(); gc.close()


@pytest.mark.parametrize("kord_tm,kord,last,tracer,hydro",
                         [(-9, 9, 0, 1, 0), (-10, 10, 1, 0, 0), (9, 8, 0, 1, 0), (-9, 11, 1, 1, 0), (10, 12, 0, 0, 0), (-13, 13, 0, 0, 0),
                          (-9, 9, 0, 1, 1), (-8, 8, 1, 0, 1), (10, 10, 0, 0, 1)])
def test_lagrangian_to_eulerian_matches_the_oracle(kord_tm, kord, last, tracer, hydro):
    case, oc, gc = _pair(hydrostatic=hydro)
    reg = _regions(case.bounds, bool(hydro))
    for t in oc.tiles:
        for e in (oc.eng[t], gc.eng[t]):
            e.call("lagrangian_to_eulerian", last, kord, kord, kord_tm, tracer, kord)
        _assert(H.compare(oc.eng[t], gc.eng[t], reg), TOL)
    oc.close(); gc.close()


def test_unsupported_remap_options_are_errors():
    case = H.Case(12, 8, "A", state="baroclinic")
    gc = H.CudaCube(case)
    e = gc.eng[1]
    for args in [(0, 0, 9, -9, 0, 9), (0, 9, -9, -9, 0, 9), (0, 9, 9, -16, 0, 9), (0, 9, 9, -9, 1, 16)]:
        with pytest.raises(RuntimeError, match="remap"):
            e.call("lagrangian_to_eulerian", *args)
    gc.close()


def test_k_split_loop_dyn_core_and_remap():
    """fv_dynamics' k_split loop (fv_dynamics.F90:478-625) with device-resident state: dyn_core -> remap (not last) -> dyn_core ->
    remap (last step), both sides driven the same way.  Everything but w stays within 1e-9; w -- a residual of nearly balanced
    forces, max |w| ~ 0.03 m/s -- differs by 2.2e-7 of its maximum (7e-9 m/s): the 1e-11 differences the first dyn_core leaves
    reach the discontinuous switches of the remap (the 2-delta-z / extremum tests of cs_profile), one of which flips in a few
    columns; same bound as the 8-substep run of tests/test_gpu_parity.py."""
    case = H.Case(N, NPZ, "A", state="baroclinic")
    oc, gc = H.OracleCube(case), H.CudaCube(case)
    for last in (0, 1):
        oc.dyn_core(900.0, 2)
        gc.dyn_core(900.0, 2)
        for t in oc.tiles:
            for e in (oc.eng[t], gc.eng[t]):
                e.call("lagrangian_to_eulerian", last, 9, 9, -9, 0, 9)
    reg = _regions(case.bounds)
    reg.pop("WORK_Q"); reg.pop("OMGA")
    for t in oc.tiles:
        res = H.compare(oc.eng[t], gc.eng[t], reg)
        _assert({k: v for k, v in res.items() if k != "W"}, 1e-9)
        _assert({"W": res["W"]}, 1e-5)
    u = gc.eng[1].get("U")
    assert np.isfinite(u).all() and np.abs(u).max() < 60.0
    oc.close(); gc.close()


@pytest.mark.parametrize("hord_tr", [0, 8])
def test_fv_dynamics_step_matches_the_oracle(hord_tr):
    """fv3_fv_dynamics (fv_dynamics.F90:303-662, dry adiabatic subset): T -> theta_v, two k_split iterations of {dyn_core with two
    acoustic substeps, tracer_2d, vertical remap}, omega filter, T out -- one C call against the same sequence of oracle stages."""
    case = H.Case(N, NPZ, "A", state="baroclinic")
    oc, gc = H.OracleCube(case), H.CudaCube(case)
    rng = np.random.default_rng(5)
    kappa = case.consts["kappa"]
    for t in oc.tiles:        # the initial state carries theta: make it a temperature, as fv_dynamics expects on entry
        eo = oc.eng[t]
        pt, delp = eo.get("PT"), eo.get("DELP")
        p = case.ak[0] + np.cumsum(delp, axis=0) - 0.5 * delp
        eo.put("PT", pt * p ** kappa)
        q = eo.get("WORK_Q"); q[...] = rng.uniform(0.0, 1.0, q.shape); eo.put("WORK_Q", q)
        for f in ("PT", "WORK_Q"):
            gc.eng[t].put(f, eo.get(f))
    oc.fv_dynamics(1800.0, 2, 2, 9, 9, -9, 9, hord_tr, 1)
    gc.fv_dynamics(1800.0, 2, 2, 9, 9, -9, 9, hord_tr, 1)
    reg = _regions(case.bounds)
    if not hord_tr:
        reg.pop("WORK_Q")
    for t in oc.tiles:
        res = H.compare(oc.eng[t], gc.eng[t], reg)
        _assert({k: v for k, v in res.items() if k not in ("W", "OMGA")}, 1e-9)
        _assert({k: res[k] for k in ("W", "OMGA")}, 1e-5)      # (see test_k_split_loop_dyn_core_and_remap)
    T = H.sub(gc.eng[1], "PT", gc.eng[1].get("PT"), 1, N, 1, N)
    assert 150.0 < T.min() and T.max() < 350.0                  # a temperature again
    oc.close(); gc.close()


@pytest.mark.parametrize("kord", [8, 10, 11, 13, 14])
def test_remap_known_answer_linear_profile_on_the_gpu(kord):
    """KNOWN ANSWER on the device, no oracle involved: a profile linear in pressure must be remapped to the analytic layer means of
    the new levels by every scheme (exact interface values, exact parabola, limiters inactive on a monotone profile)."""
    case = H.Case(N, NPZ, "A", state="baroclinic")
    gc = H.CudaCube(case)
    gc.dyn_core(900.0, 2)                                   # deforms the Lagrangian surfaces and leaves pe current
    for t in (1, 4):
        e = gc.eng[t]
        pe = np.transpose(H.sub(e, "PE", e.get("PE"), 1, N, 1, N), (1, 0, 2))       # (k, j, i)
        p2 = case.ak[:, None, None] + case.bk[:, None, None] * pe[-1][None]
        p2[0] = case.ak[0]; p2[-1] = pe[-1]
        assert np.abs(np.diff(pe, axis=0) - np.diff(p2, axis=0)).max() > 1e-3
        mean = lambda p: 2.0 + 3.0e-5 * 0.5 * (p[:-1] + p[1:])
        full = e.get("WORK_Q")
        H.sub(e, "WORK_Q", full, 1, N, 1, N)[...] = mean(pe)
        e.put("WORK_Q", full)
        for mode, iv in ((0, 1), (1, 1), (2, 0)):
            q0 = e.get("WORK_Q")
            e.call("remap_work_q", mode, iv, kord, 0.0)
            out = H.sub(e, "WORK_Q", e.get("WORK_Q"), 1, N, 1, N)
            assert np.abs(out - mean(p2)).max() / np.abs(mean(p2)).max() < 1e-12, (t, mode)
            e.put("WORK_Q", q0)
    gc.close()


@pytest.mark.parametrize("nq", [3, 6])
def test_fv_dynamics_with_several_tracers(nq):
    """nq tracers through fv3_fv_dynamics: tracer_2d advects all of them with one CFL / sub-cycle count and one dp1 -> dp2 per
    sub-cycle (fv_tracer2d.F90:206-275), the remap maps them with map1_q2 (nq <= 5) or in the operation order of mapn_tracer
    (nq > 5: fv_mapz.F90:390-408).  Against the oracle; a tracer that is constant must stay constant, and the tracer buffers must
    come back in their slots (the sub-cycles rotate them through a spare buffer)."""
    case = H.Case(N, NPZ, "A", state="baroclinic")
    oc, gc = H.OracleCube(case), H.CudaCube(case)
    oc.set_num_tracers(nq); gc.set_num_tracers(nq)
    rng = np.random.default_rng(3)
    kappa = case.consts["kappa"]
    for t in oc.tiles:
        eo, eg = oc.eng[t], gc.eng[t]
        pt, delp = eo.get("PT"), eo.get("DELP")
        p = case.ak[0] + np.cumsum(delp, axis=0) - 0.5 * delp
        eo.put("PT", pt * p ** kappa); eg.put("PT", eo.get("PT"))
        for iq in range(nq):
            q = np.full(eo.shape("WORK_Q"), 0.25 * (iq + 1)) if iq == 1 else rng.uniform(0.0, 1.0, eo.shape("WORK_Q")) * (iq + 1)
            for e in (eo, eg):
                e.call("select_tracer", iq); e.put("WORK_Q", q)
        for e in (eo, eg):
            e.call("select_tracer", 0)
    oc.fv_dynamics(1800.0, 2, 2, 9, 9, -9, 9, 8, 1)
    gc.fv_dynamics(1800.0, 2, 2, 9, 9, -9, 9, 8, 1)
    b = case.bounds
    reg = {"WORK_Q": (b["is_"], b["ie"], b["js"], b["je"])}
    for iq in range(nq):
        oc.select_tracer(iq); gc.select_tracer(iq)
        for t in oc.tiles:
            _assert(H.compare(oc.eng[t], gc.eng[t], reg), 1e-9)
        if iq == 1:
            q = H.sub(gc.eng[2], "WORK_Q", gc.eng[2].get("WORK_Q"), 1, N, 1, N)
            assert np.abs(q - 0.5).max() < 1e-12
    reg2 = _regions(case.bounds); reg2.pop("WORK_Q")
    for t in oc.tiles:
        res = H.compare(oc.eng[t], gc.eng[t], reg2)
        _assert({k: v for k, v in res.items() if k not in ("W", "OMGA")}, 1e-9)
    oc.close(); gc.close()


def test_fv_dynamics_with_water_vapour():
    """fv3_fv_dynamics_qv: tracer 0 is the specific humidity (no condensates): theta_v = T (1 + zvir q_v) / pkz on entry
    (fv_dynamics.F90:303-398), T = T_v / (1 + zvir q_v) after the last remap (fv_mapz.F90:792-822).  Against the oracle, and the
    moisture must matter (a moist run differs from the dry one)."""
    zvir = 461.5 / 287.05 - 1.0
    case = H.Case(N, NPZ, "A", state="baroclinic")
    oc, gc, gd = H.OracleCube(case), H.CudaCube(case), H.CudaCube(case)
    for cube in (oc, gc, gd):
        cube.set_num_tracers(2)
    rng = np.random.default_rng(9)
    kappa = case.consts["kappa"]
    for t in oc.tiles:
        eo = oc.eng[t]
        pt, delp = eo.get("PT"), eo.get("DELP")
        p = case.ak[0] + np.cumsum(delp, axis=0) - 0.5 * delp
        T0 = pt * p ** kappa
        qv = 0.015 * (p / 1.0e5) ** 3 * rng.uniform(0.5, 1.0, pt.shape)          # a few g/kg near the surface
        q1 = rng.uniform(0.0, 1.0, pt.shape)
        for e in (eo, gc.eng[t], gd.eng[t]):
            e.put("PT", T0)
            e.call("select_tracer", 0); e.put("WORK_Q", qv)
            e.call("select_tracer", 1); e.put("WORK_Q", q1)
            e.call("select_tracer", 0)
    oc.fv_dynamics(1800.0, 2, 2, 9, 9, -9, 9, 8, 1, sphum=0, zvir=zvir)
    gc.fv_dynamics(1800.0, 2, 2, 9, 9, -9, 9, 8, 1, sphum=0, zvir=zvir)
    gd.fv_dynamics(1800.0, 2, 2, 9, 9, -9, 9, 8, 1)
    reg = _regions(case.bounds)
    for t in oc.tiles:
        res = H.compare(oc.eng[t], gc.eng[t], reg)
        _assert({k: v for k, v in res.items() if k not in ("W", "OMGA")}, 1e-9)
    Tm = H.sub(gc.eng[1], "PT", gc.eng[1].get("PT"), 1, N, 1, N); Td = H.sub(gd.eng[1], "PT", gd.eng[1].get("PT"), 1, N, 1, N)
    assert 150.0 < Tm.min() and Tm.max() < 350.0
    assert np.abs(Tm - Td).max() > 1e-3
    oc.close(); gc.close(); gd.close()
