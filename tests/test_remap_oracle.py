"""Oracle-side checks of the vertical remap (oracle/remap.cpp: fv_mapz.F90 Lagrangian_to_Eulerian and the fv_operators.F90 column
operators).  The reference has no test or golden vector for these routines ("parity unpinned"), so the restatement is held to
the properties the algorithm defines:
  * a remap onto the SAME levels returns the layer means (every new layer lies within one old layer: the parabola integrates
    to its mean), for every scheme 3..15 and every boundary mode;
  * the mapping is conservative: sum_k q dp is unchanged for map1_ppm / map_scalar / map1_q2;
  * a constant stays constant, a positive tracer stays non-negative (iv = 0);
  * Lagrangian_to_Eulerian leaves the surface pressure, the column height and the column integral of w unchanged, puts delp on
    the hybrid levels, and returns pe / peln / pk / pkz consistent with each other.
"""
import ctypes as C

import numpy as np
import pytest

import harness as H

KORDS = [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]   # <= 7: ppm_profile (1, 2 = 3: the same limiter), > 7: cs / scalar_profile
N, NPZ = 12, 16


def _cube(substeps=2, **over):
    case = H.Case(N, NPZ, "A", state="baroclinic", flags_override=over or None)
    oc = H.OracleCube(case)
    if substeps:
        oc.dyn_core(600.0 * substeps, substeps)
    return case, oc


def _sec(e, name, extra_i=0, extra_j=0):
    return H.sub(e, name, e.get(name), 1, N + extra_i, 1, N + extra_j).copy()


def _pe(e):
    """pe on the compute domain as (k, j, i)"""
    pe = H.sub(e, "PE", e.get("PE"), 1, N, 1, N)          # (j, k, i)
    return np.transpose(pe, (1, 0, 2)).copy()


def _hybrid(case, pe):
    ak, bk = case.ak, case.bk
    p2 = ak[:, None, None] + bk[:, None, None] * pe[-1][None]
    p2[0] = case.ak[0]; p2[-1] = pe[-1]
    return p2


def _set_pe_from_delp(case, e):
    """pe(is-1:ie+1, 1:km+1, js-1:je+1) = ptop + the partial sums of delp (what the vertical solver / pe_halo leave behind)"""
    delp = H.sub(e, "DELP", e.get("DELP"), 0, N + 1, 0, N + 1)               # (k, j, i)
    pe = np.concatenate([np.full((1,) + delp.shape[1:], case.ak[0]), case.ak[0] + np.cumsum(delp, axis=0)], axis=0)
    e.put("PE", np.ascontiguousarray(np.transpose(pe, (1, 0, 2))))          # host layout: k in the middle


def _set_q(e, q):
    full = e.get("WORK_Q")
    H.sub(e, "WORK_Q", full, 1, N, 1, N)[...] = q
    e.put("WORK_Q", full)


@pytest.mark.parametrize("kord", KORDS)
@pytest.mark.parametrize("mode,iv", [(0, 1), (1, 1), (1, -1), (1, -2), (2, 0)])
def test_remap_onto_the_same_levels_is_the_identity(kord, mode, iv):
    case, oc = _cube(substeps=0)           # initial state: delp sits exactly on the hybrid levels
    e = oc.eng[1]
    _set_pe_from_delp(case, e)
    rng = np.random.default_rng(kord * 10 + mode)
    q = rng.uniform(0.5, 1.5, (NPZ, N, N)) * (1.0 + np.linspace(0, 1, NPZ))[:, None, None]
    _set_q(e, q)
    if iv == -2:
        ws = e.get("WS"); ws[...] = 0.3; e.put("WS", ws)
    e.call("remap_work_q", mode, iv, kord, 0.0)
    out = _sec(e, "WORK_Q")
    assert np.abs(out - q).max() / np.abs(q).max() < 2e-13
    oc.close()


@pytest.mark.parametrize("kord", KORDS)
@pytest.mark.parametrize("mode,iv", [(0, 1), (1, 1), (1, -1), (1, -2), (2, 0)])
def test_remap_is_conservative_and_keeps_a_constant(kord, mode, iv):
    case, oc = _cube(substeps=2)           # two acoustic substeps have deformed the Lagrangian surfaces
    e = oc.eng[3]
    pe = _pe(e)
    p2 = _hybrid(case, pe)
    dp1, dp2 = np.diff(pe, axis=0), np.diff(p2, axis=0)
    assert np.abs(dp1 - dp2).max() > 1e-3, "the test needs deformed levels"
    rng = np.random.default_rng(7)
    q = np.abs(H.sub(e, "PT", e.get("PT"), 1, N, 1, N)) * rng.uniform(0.9, 1.1, (NPZ, N, N))
    _set_q(e, q)
    e.call("remap_work_q", mode, iv, kord, 0.0)
    out = _sec(e, "WORK_Q")
    s1, s2 = (q * dp1).sum(0), (out * dp2).sum(0)
    assert np.abs(s2 - s1).max() / np.abs(s1).max() < 1e-13
    if iv != -2:                            # (the lower boundary condition of iv = -2 pulls the bottom layer towards ws)
        _set_q(e, np.full((NPZ, N, N), 3.25))
        e.call("remap_work_q", mode, iv, kord, 0.0)
        assert np.abs(_sec(e, "WORK_Q") - 3.25).max() < 1e-12
    oc.close()


@pytest.mark.parametrize("kord", [k for k in KORDS if k != 6])   # (6 is the unlimited parabola of ppm_profile: no positivity promise)
def test_tracer_remap_keeps_a_positive_field_non_negative(kord):
    case, oc = _cube(substeps=2)
    e = oc.eng[2]
    q = np.zeros((NPZ, N, N))
    q[5:8] = 1.0; q[11] = 1e-3                      # sharp layers: the unlimited parabolas undershoot
    _set_q(e, q)
    e.call("remap_work_q", 2, 0, kord, 0.0)
    out = _sec(e, "WORK_Q")
    assert out.min() >= 0.0               # (the schemes 8..13 are not strictly monotone: a top hat may overshoot by ~1e-4)
    assert out.max() <= 1.01
    oc.close()


def test_unsupported_schemes_are_errors():
    case, oc = _cube(substeps=0)
    e = oc.eng[1]
    for kord in (0, 16, 17):
        with pytest.raises(RuntimeError):
            e.call("remap_work_q", 0, 1, kord, 0.0)
    with pytest.raises(RuntimeError):
        e.call("lagrangian_to_eulerian", 0, 9, -9, -9, 0, 9)     # kord_wz < 0: the iv = -3 branch
    oc.close()


@pytest.mark.parametrize("kord_tm,last,hydro", [(-9, 0, 0), (-10, 1, 0), (9, 0, 0), (-9, 0, 1), (-8, 1, 1), (-7, 0, 0), (4, 1, 0),
                                                  (-6, 1, 1)])
def test_lagrangian_to_eulerian_invariants(kord_tm, last, hydro):
    case, oc = _cube(substeps=2, hydrostatic=hydro)
    f = case.flags
    kappa, rdgas, grav = case.consts["kappa"], case.consts["rdgas"], case.consts["grav"]
    for t in (1, 4):
        e = oc.eng[t]
        pe = _pe(e)
        w0, dz0, dp0 = _sec(e, "W"), _sec(e, "DELZ"), _sec(e, "DELP")
        pt0 = _sec(e, "PT")
        om0 = _sec(e, "OMGA")
        u0 = _sec(e, "U", 0, 1)
        ko = 9 if abs(kord_tm) > 7 else abs(kord_tm)          # (the ppm_profile cases remap the winds and w with it as well)
        e.call("lagrangian_to_eulerian", last, ko, ko, kord_tm, 0, 9)
        p2 = _hybrid(case, pe)
        dp1 = _sec(e, "DELP")
        assert np.abs(dp1 - np.diff(p2, axis=0)).max() / dp1.max() < 1e-14          # delp on the hybrid levels
        assert np.abs(_pe(e) - p2).max() / p2.max() < 1e-14                          # pe follows
        peln = np.transpose(H.sub(e, "PELN", e.get("PELN"), 1, N, 1, N), (1, 0, 2))
        assert np.abs(peln - np.log(p2)).max() < 1e-12
        pk = _sec(e, "PK")
        assert np.abs(pk - np.exp(kappa * np.log(p2))).max() / pk.max() < 1e-12
        pt1, pkz = _sec(e, "PT"), _sec(e, "PKZ")
        if not hydro:
            w1, dz1 = _sec(e, "W"), _sec(e, "DELZ")
            assert np.abs(dz1.sum(0) - dz0.sum(0)).max() / np.abs(dz0.sum(0)).max() < 1e-13   # the column keeps its height
            assert np.abs((w1 * dp1).sum(0) - (w0 * dp0).sum(0)).max() / np.abs(w0 * dp0).sum(0).max() < 1e-12
            tv = pt1 if last else pt1 * pkz                                          # T_v: what pkz was formed from
            ex = kappa if kord_tm < 0 else rdgas / (case.consts["cp_air"] - rdgas)
            ref = np.exp(ex * np.log(-rdgas / grav * dp1 / dz1 * (tv if kord_tm < 0 else tv / pkz)))
            assert np.abs(pkz - ref).max() / pkz.max() < 1e-12
        else:
            assert np.abs(pkz - np.diff(pk, axis=0) / (kappa * np.diff(peln, axis=0))).max() / pkz.max() < 1e-13
        # theta_v changed little (the levels moved by a fraction of a layer); last_step returns T_v instead
        th1 = pt1 / pkz if last else pt1
        assert np.abs(th1 - pt0).max() / np.abs(pt0).max() < 2e-2
        if last:
            om1 = _sec(e, "OMGA")
            assert om1.min() >= min(om0.min(), 0.0) - 1e-12 and om1.max() <= max(om0.max(), 0.0) + 1e-12   # linear interpolation
        u1 = _sec(e, "U", 0, 1)
        assert np.abs(u1 - u0).max() < 0.5 and np.isfinite(u1).all()
    oc.close()


def _analytic_case(e, case, coef):
    """q(p) = c0 + c1 p + c2 p^2 as exact layer means on the deformed levels of engine e; returns (q, expected means on the
    hybrid target levels)"""
    pe = _pe(e)
    p2 = _hybrid(case, pe)
    c0, c1, c2 = coef

    def mean(p):
        a, b = p[:-1], p[1:]
        return c0 + c1 * 0.5 * (a + b) + c2 * (a * a + a * b + b * b) / 3.0
    return mean(pe), mean(p2)


@pytest.mark.parametrize("kord", KORDS)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_remap_known_answer_linear_profile(kord, mode):
    """KNOWN ANSWER (independent of any restatement): a profile linear in pressure is reproduced exactly by the 4th-order interface
    values and the parabolic sub-grid profile, and its monotonicity keeps every limiter of every scheme inactive, so the remapped
    layer means must equal the analytic means on the new levels -- in the interior; the top / bottom two layers use one-sided
    closures that are exact for a linear profile as well."""
    case, oc = _cube(substeps=2)
    e = oc.eng[5]
    q, want = _analytic_case(e, case, (2.0, 3.0e-5, 0.0))
    _set_q(e, q)
    e.call("remap_work_q", mode, 0 if mode == 2 else 1, kord, 0.0)
    out = _sec(e, "WORK_Q")
    assert np.abs(out - want).max() / np.abs(want).max() < 1e-12
    oc.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_remap_known_answer_quadratic_profile(mode):
    """KNOWN ANSWER: with the unlimited scheme (kord 13) a profile quadratic in pressure is remapped exactly (the interface solver
    is exact for cubics, the parabola for quadratics) away from the bottom, whose closure is second order: its error (6e-7 here)
    decays by ~4 per level through the tridiagonal solve, so the upper levels see the exact answer."""
    case, oc = _cube(substeps=2)
    e = oc.eng[2]
    q, want = _analytic_case(e, case, (1.0, 2.0e-5, 1.5e-10))
    _set_q(e, q)
    e.call("remap_work_q", mode, 1, 13, 0.0)
    out = _sec(e, "WORK_Q")
    err = np.abs(out - want) / np.abs(want).max()
    lev = err.max(axis=(1, 2))
    assert lev[:6].max() < 5e-11, lev
    assert lev.max() < 1e-5, lev
    assert all(lev[k] < 0.6 * lev[k + 1] for k in range(9, 13)), lev          # geometric decay away from the bottom closure
    oc.close()


@pytest.mark.parametrize("kord", [3, 4, 6, 7])
@pytest.mark.parametrize("mode", [0, 1])
def test_ppm_profile_known_answer_quadratic_profile(kord, mode):
    """KNOWN ANSWER for ppm_profile (kord <= 7): its interface values come from an explicit local 4th-order formula, so -- unlike the
    tridiagonal solve of the cs profiles -- the one-sided closures of the top / bottom layers do not leak into the interior: a
    profile quadratic and monotone in pressure (no limiter active) is remapped to rounding error away from the three boundary
    layers on each side."""
    case, oc = _cube(substeps=2)
    e = oc.eng[2]
    q, want = _analytic_case(e, case, (1.0, 2.0e-5, 1.5e-10))
    _set_q(e, q)
    e.call("remap_work_q", mode, 1, kord, 0.0)
    out = _sec(e, "WORK_Q")
    lev = (np.abs(out - want) / np.abs(want).max()).max(axis=(1, 2))
    assert lev[3:NPZ - 3].max() < 2e-15, lev
    assert lev.max() < 1e-5, lev
    oc.close()


def test_fillz_properties_and_known_answer():
    """fillz (fv_fill.F90:34-139): a non-negative tracer is left untouched bit for bit; negative layers borrow from their vertical
    neighbours, so the column integral sum_k q dp is kept and -- wherever the column total is positive -- the result is
    non-negative; a hand-computed column as the known answer."""
    case, oc = _cube(substeps=0)
    e = oc.eng[3]
    dp = _sec(e, "DELP")
    rng = np.random.default_rng(5)
    q = rng.uniform(0.0, 1.0, (NPZ, N, N))
    _set_q(e, q)
    e.call("fillz")
    assert np.array_equal(_sec(e, "WORK_Q"), q)
    q = rng.uniform(-0.3, 1.0, (NPZ, N, N)) * rng.integers(0, 2, (NPZ, N, N))          # negative layers next to empty and full ones
    q[:, 0, 0] = -np.abs(q[:, 0, 0])                                                    # one column without any mass to borrow
    _set_q(e, q)
    e.call("fillz")
    out = _sec(e, "WORK_Q")
    s0, s1 = (q * dp).sum(0), (out * dp).sum(0)
    assert np.abs(s1 - s0).max() / np.abs(s0).max() < 1e-14
    fixable = (q[1:] * dp[1:]).sum(0) > 0.0                                             # (the non-local fix works on layers 2..km)
    assert fixable.sum() > 100 and not fixable[0, 0]
    assert out[1:][:, fixable].min() >= 0.0
    assert out[0].min() > -1e-16            # (layer 1 lends to layer 2 after its own test: q1 - (q1 dp1) / dp1 may round to -1 ulp)
    assert np.array_equal(out[:, 0, 0], np.concatenate([[0.0], q[1:, 0, 0]]) + np.concatenate([[0.0, q[0, 0, 0] * dp[0, 0, 0] / dp[1, 0, 0]], np.zeros(NPZ - 2)]))
    # known answer on layers of equal thickness: layer 3 borrows 0.5 from layer 2 (above) and the rest from layer 4 (below)
    full = e.get("DELP"); full[...] = 100.0; e.put("DELP", full)
    col = np.array([1.0, 0.5, -1.0, 2.0] + [0.25] * (NPZ - 4))
    _set_q(e, np.broadcast_to(col[:, None, None], (NPZ, N, N)).copy())
    e.call("fillz")
    want = np.array([1.0, 0.0, 0.0, 1.5] + [0.25] * (NPZ - 4))
    assert np.abs(_sec(e, "WORK_Q") - want[:, None, None]).max() < 1e-15
    oc.close()


def test_lagrangian_to_eulerian_with_fill_keeps_the_tracer_mass_and_removes_negatives():
    case, oc = _cube(substeps=2)
    e = oc.eng[2]
    q = np.zeros((NPZ, N, N)); q[4:7] = 1.0; q[10] = 1e-3
    dp0 = _sec(e, "DELP")
    outs = {}
    for fill in (0, 1):
        oc2_case, oc2 = _cube(substeps=2)
        e2 = oc2.eng[2]
        _set_q(e2, q)
        e2.call("set_tracer_fill", fill)
        e2.call("lagrangian_to_eulerian", 0, 9, 9, -9, 1, 6)       # kord_tr = 6: the unlimited parabola undershoots
        outs[fill] = (_sec(e2, "WORK_Q"), _sec(e2, "DELP"))
        oc2.close()
    (a, dpa), (b, dpb) = outs[0], outs[1]
    assert a.min() < -1e-6, "the test needs undershoots to fill"
    assert b.min() > -1e-16
    assert np.abs((a * dpa).sum(0) - (b * dpb).sum(0)).max() / (a * dpa).sum(0).max() < 1e-13
    assert np.abs((b * dpb).sum(0) - (q * dp0).sum(0)).max() / (q * dp0).sum(0).max() < 1e-13
    oc.close()


def test_more_than_five_tracers_with_kord_tr_below_8_follow_mapn_tracer():
    """nq > 5: fv_mapz calls mapn_tracer, which has no ppm_profile branch -- scalar_profile runs whatever kord_tr is and treats
    abs(kord) 0..8 alike (fv_operators.F90:262-273, 756): kord_tr = 5 must give what kord_tr = 8 gives, and not what the
    ppm_profile of map1_q2 (nq <= 5) gives."""
    q = np.zeros((NPZ, N, N)); q[4:7] = 1.0; q[10] = 1e-3
    res = {}
    for nq, kord_tr in [(6, 5), (6, 8), (1, 5)]:
        case, oc = _cube(substeps=2)
        e = oc.eng[2]
        e.call("set_num_tracers", nq)
        _set_q(e, q)
        e.call("lagrangian_to_eulerian", 0, 9, 9, -9, nq, kord_tr)
        res[(nq, kord_tr)] = _sec(e, "WORK_Q")
        oc.close()
    assert np.array_equal(res[(6, 5)], res[(6, 8)])
    assert np.abs(res[(6, 5)] - res[(1, 5)]).max() > 1e-6
