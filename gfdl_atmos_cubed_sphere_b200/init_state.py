"""Idealised initial states for the hot-path harness (host-side fixture generator).

* ``baroclinic_wave``: the Jablonowski-Williamson (2006) steady state + wind perturbation that
  ``tools/test_cases.F90:1575-1890`` (test_case 12/13) initialises, evaluated pointwise from
  the published analytic formulas at the D-grid edge mid-points / cell centres (the reference
  additionally uses Gaussian quadrature along edges, ``test_cases.F90:1650-1700``; the
  difference is O(dx^2) and irrelevant for oracle-vs-CUDA parity, which runs both from the
  same arrays).  Dry, ``w`` = analytic perturbation of SURVEY 8(d), ``delz`` hydrostatic
  (``init_hydro.F90`` ``p_var``: ``delz = -(rdgas/grav) T dln p``).
* ``hybrid_levels``: a smooth hybrid sigma-p coordinate with ``npz`` layers.  The reference's
  L79 table comes from the ``var_hi`` generator (``fv_eta.F90:656-666``), which is not restated
  -- stated in DESIGN.md.
* ``smooth_state``: cheap analytic fields for operator-level parity tests.

All halos are filled with the 6-tile exchange of :mod:`cubed_sphere`.
"""
from __future__ import annotations

import numpy as np

from . import cubed_sphere as cs
from .grid import latlon2xyz, CONSTANTS


def hybrid_levels(npz, ptop=100.0, p0=1.0e5, eta_c=0.15, r=1.6):
    """ak, bk (npz+1): log-stretched in pressure aloft, sigma-like near the surface."""
    s = np.linspace(0.0, 1.0, npz + 1)
    # blend log-spacing (top) with linear spacing (bottom) for well-resolved boundary layer
    lo = np.log(ptop / p0)
    eta = np.exp(lo * (1.0 - s) ** 1.4) * (1.0 - 0.35 * s * (1.0 - s))
    eta[0], eta[-1] = ptop / p0, 1.0
    eta = np.maximum.accumulate(eta)
    bk = (np.maximum(eta - eta_c, 0.0) / (1.0 - eta_c)) ** r
    ak = p0 * (eta - bk)
    ak[0], bk[0] = ptop, 0.0
    ak[-1], bk[-1] = 0.0, 1.0
    assert np.all(np.diff(ak + bk * p0) > 0.0) and np.all(np.diff(ak + bk * 5.0e4) > 0.0)
    return ak, bk


def set_eta_var_hi(km, ptop=1.0, pint=100.0e2, s_rate=1.03, rdgas=287.05, grav=9.80665):
    """ak, bk, ks of the reference's automatic hybrid-level generator var_hi (tools/fv_eta.F90:1166-1341, the non-HIWPP,
    UKMO-hybrid branch) as set_eta selects it for km = 79: ptop = 1 Pa, stretch_fac = 1.03, pint = 100 hPa
    (fv_eta.F90:656-666, :296-297).  Statement-for-statement restatement incl. sm1_edge (:2313-2346, one pass)."""
    assert km >= 26, "var_hi needs km - k_inc - 2 >= 9"
    p00, k_inc, s0, t0 = 1.0e5, 15, 0.10, 270.0
    pe1 = np.zeros(km + 2); peln = np.zeros(km + 2); ze = np.zeros(km + 2)      # 1-based
    s_fac = np.zeros(km + 1); dz = np.zeros(km + 1); dlnp = np.zeros(km + 1)
    pe1[1] = ptop; peln[1] = np.log(pe1[1]); pe1[km + 1] = p00; peln[km + 1] = np.log(pe1[km + 1])
    ztop = rdgas / grav * t0 * (peln[km + 1] - peln[1])
    s_inc = (1.0 - s0) / float(k_inc)
    s_fac[km] = s0
    for k in range(km - 1, km - k_inc - 1, -1):
        s_fac[k] = s_fac[k + 1] + s_inc
    s_fac[km - k_inc - 1] = 0.5 * (s_fac[km - k_inc] + s_rate)
    for k in range(km - k_inc - 2, 8, -1):
        s_fac[k] = s_rate * s_fac[k + 1]
    s_fac[8] = 0.5 * (1.1 + s_rate) * s_fac[9]
    s_fac[7] = 1.1 * s_fac[8]; s_fac[6] = 1.15 * s_fac[7]; s_fac[5] = 1.2 * s_fac[6]; s_fac[4] = 1.3 * s_fac[5]
    s_fac[3] = 1.4 * s_fac[4]; s_fac[2] = 1.45 * s_fac[3]; s_fac[1] = 1.5 * s_fac[2]
    sum1 = 0.0
    for k in range(1, km + 1):
        sum1 = sum1 + s_fac[k]
    dz0 = ztop / sum1
    for k in range(1, km + 1):
        dz[k] = s_fac[k] * dz0
    ze[km + 1] = 0.0
    for k in range(km, 0, -1):
        ze[k] = ze[k + 1] + dz[k]
    for k in range(1, km + 1):                     # re-scale dz with the stretched ztop
        dz[k] = dz[k] * (ztop / ze[1])
    for k in range(km, 0, -1):
        ze[k] = ze[k + 1] + dz[k]
    # sm1_edge(..., ze, ntimes = 1): note its own dz = ze(k+1) - ze(k) (negative thicknesses)
    df, k2, ntimes = 0.25, km - 1, 1
    dzs = np.zeros(km + 1); flux = np.zeros(km + 2)
    for k in range(1, km + 1):
        dzs[k] = ze[k + 1] - ze[k]
    for n in range(1, ntimes + 1):
        k1 = 2 + (ntimes - n)
        flux[k1] = 0.0; flux[k2 + 1] = 0.0
        for k in range(k1 + 1, k2 + 1):
            flux[k] = df * (dzs[k] - dzs[k - 1])
        for k in range(k1, k2 + 1):
            dzs[k] = dzs[k] - flux[k] + flux[k + 1]
    for k in range(km, 0, -1):
        ze[k] = ze[k + 1] - dzs[k]
    for k in range(1, km + 1):                     # given z --> p
        dz[k] = ze[k] - ze[k + 1]
        dlnp[k] = grav * dz[k] / (rdgas * t0)
    for k in range(2, km + 1):
        peln[k] = peln[k - 1] + dlnp[k - 1]
        pe1[k] = np.exp(peln[k])
    ks = 0
    for k in range(2, km + 1):
        if pint < pe1[k]:
            ks = k - 1
            break
    eta = np.zeros(km + 2)
    for k in range(1, km + 2):
        eta[k] = pe1[k] / pe1[km + 1]
    ep, es = eta[ks + 1], eta[km]
    alpha = (ep ** 2 - 2.0 * ep * es) / (es - ep) ** 2
    beta = 2.0 * ep * es ** 2 / (es - ep) ** 2
    gama = -(ep * es) ** 2 / (es - ep) ** 2
    ak = np.zeros(km + 2); bk = np.zeros(km + 2)
    for k in range(1, ks + 2):
        ak[k] = eta[k] * 1.0e5; bk[k] = 0.0
    for k in range(ks + 2, km + 1):
        ak[k] = alpha * eta[k] + beta + gama / eta[k]
        ak[k] = ak[k] * 1.0e5
    ak[km + 1] = 0.0
    for k in range(ks + 2, km + 1):
        bk[k] = (pe1[k] - ak[k]) / pe1[km + 1]
    bk[km + 1] = 1.0
    return ak[1:].copy(), bk[1:].copy(), ks


def set_eta_var_gfs(km, ptop=1.0, pint=75.0e2, s_rate=1.028, rdgas=287.05, grav=9.80665):
    """ak, bk, ks of the reference's generator var_gfs (tools/fv_eta.F90:1002-1164, UKMO-hybrid branch) as set_eta selects it for
    km = 127, default npz_type (BASELINE config 5: ptop = 1 Pa, pint = 75 hPa, stretch_fac = 1.028; fv_eta.F90:726-744).
    Differs from var_hi in the tunables (k_inc = 25, s0 = 0.13), the top-layer factors, and in having no sm1_edge pass."""
    assert km >= 36, "var_gfs needs km - k_inc - 1 >= 9"
    p00, k_inc, s0, t0 = 1.0e5, 25, 0.13, 270.0
    pe1 = np.zeros(km + 2); peln = np.zeros(km + 2); ze = np.zeros(km + 2)      # 1-based
    s_fac = np.zeros(km + 1); dz = np.zeros(km + 1); dlnp = np.zeros(km + 1)
    pe1[1] = ptop; peln[1] = np.log(pe1[1]); pe1[km + 1] = p00; peln[km + 1] = np.log(pe1[km + 1])
    ztop = rdgas / grav * t0 * (peln[km + 1] - peln[1])
    s_inc = (1.0 - s0) / float(k_inc)
    s_fac[km] = s0
    for k in range(km - 1, km - k_inc - 1, -1):
        s_fac[k] = s_fac[k + 1] + s_inc
    for k in range(km - k_inc - 1, 8, -1):
        s_fac[k] = s_rate * s_fac[k + 1]
    s_fac[8] = 0.5 * (1.1 + s_rate) * s_fac[9]
    s_fac[7] = 1.10 * s_fac[8]; s_fac[6] = 1.15 * s_fac[7]; s_fac[5] = 1.20 * s_fac[6]; s_fac[4] = 1.26 * s_fac[5]
    s_fac[3] = 1.33 * s_fac[4]; s_fac[2] = 1.41 * s_fac[3]; s_fac[1] = 1.60 * s_fac[2]
    sum1 = 0.0
    for k in range(1, km + 1):
        sum1 = sum1 + s_fac[k]
    dz0 = ztop / sum1
    for k in range(1, km + 1):
        dz[k] = s_fac[k] * dz0
    ze[km + 1] = 0.0
    for k in range(km, 0, -1):
        ze[k] = ze[k + 1] + dz[k]
    for k in range(1, km + 1):                     # re-scale dz with the stretched ztop
        dz[k] = dz[k] * (ztop / ze[1])
    for k in range(km, 0, -1):
        ze[k] = ze[k + 1] + dz[k]
    for k in range(1, km + 1):                     # given z --> p
        dz[k] = ze[k] - ze[k + 1]
        dlnp[k] = grav * dz[k] / (rdgas * t0)
    for k in range(2, km + 1):
        peln[k] = peln[k - 1] + dlnp[k - 1]
        pe1[k] = np.exp(peln[k])
    ks = 0
    for k in range(2, km + 1):
        if pint < pe1[k]:
            ks = k - 1
            break
    return _ukmo_hybrid(km, pe1, ks)


def _ukmo_hybrid(km, pe1, ks):
    """pe1 -> (ak, bk): pure pressure down to interface ks + 1, UKMO hybrid below (fv_eta.F90:1121-1149 = :1301-1329)"""
    eta = np.zeros(km + 2)
    for k in range(1, km + 2):
        eta[k] = pe1[k] / pe1[km + 1]
    ep, es = eta[ks + 1], eta[km]
    alpha = (ep ** 2 - 2.0 * ep * es) / (es - ep) ** 2
    beta = 2.0 * ep * es ** 2 / (es - ep) ** 2
    gama = -(ep * es) ** 2 / (es - ep) ** 2
    ak = np.zeros(km + 2); bk = np.zeros(km + 2)
    for k in range(1, ks + 2):
        ak[k] = eta[k] * 1.0e5; bk[k] = 0.0
    for k in range(ks + 2, km + 1):
        ak[k] = alpha * eta[k] + beta + gama / eta[k]
        ak[k] = ak[k] * 1.0e5
    ak[km + 1] = 0.0
    for k in range(ks + 2, km + 1):
        bk[k] = (pe1[k] - ak[k]) / pe1[km + 1]
    bk[km + 1] = 1.0
    return ak[1:].copy(), bk[1:].copy(), ks


# set_eta, km = 32, default npz_type (fv_eta.F90:410-428: ks = 7, ak = a32, bk = b32); table values from tools/fv_eta.h:114-136
_A32 = (100.00000, 400.00000, 818.60211, 1378.88653, 2091.79519, 2983.64084, 4121.78960, 5579.22148, 6907.19063, 7735.78639,
        8197.66476, 8377.95525, 8331.69594, 8094.72213, 7690.85756, 7139.01788, 6464.80251, 5712.35727, 4940.05347, 4198.60465,
        3516.63294, 2905.19863, 2366.73733, 1899.19455, 1497.78137, 1156.25252, 867.79199, 625.59324, 423.21322, 254.76613,
        115.06646, 0.00000, 0.00000)
_B32 = (0.00000, 0.00000, 0.00000, 0.00000, 0.00000, 0.00000, 0.00000, 0.00000, 0.00513, 0.01969, 0.04299, 0.07477, 0.11508,
        0.16408, 0.22198, 0.28865, 0.36281, 0.44112, 0.51882, 0.59185, 0.65810, 0.71694, 0.76843, 0.81293, 0.85100, 0.88331,
        0.91055, 0.93338, 0.95244, 0.96828, 0.98142, 0.99223, 1.00000)


def model_levels(npz):
    """Hybrid levels for a run: the reference's own set_eta result where it is restated -- npz = 79 (var_hi generator), npz = 127 (var_gfs generator: BASELINE config 5) and
    npz = 32 (the a32/b32 table: BASELINE config 1b, C48 L32) -- else the generic hybrid_levels."""
    if npz == 79:
        ak, bk, _ = set_eta_var_hi(79)
        return ak, bk
    if npz == 127:
        ak, bk, _ = set_eta_var_gfs(127)
        return ak, bk
    if npz == 32:
        return np.array(_A32, dtype=np.float64), np.array(_B32, dtype=np.float64)
    return hybrid_levels(npz)


def _edge_dirs(g):
    """Unit tangent vectors + mid-point lon/lat of D-grid u (south) and v (west) edges."""
    lon, lat = g.arr["grid"]
    P = latlon2xyz(lon, lat)                      # (nj+1, ni+1, 3)
    def mid_and_dir(a, b):
        m = a + b
        m /= np.linalg.norm(m, axis=-1, keepdims=True)
        d = b - a
        d -= np.sum(d * m, -1, keepdims=True) * m
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        return m, d
    mu, du = mid_and_dir(P[:, :-1], P[:, 1:])    # u: (nj+1, ni)
    mv, dv = mid_and_dir(P[:-1, :], P[1:, :])    # v: (nj, ni+1)
    return mu, du, mv, dv


def _wind_on_edge(m, d, ufun):
    lon = np.arctan2(m[..., 1], m[..., 0])
    lat = np.arcsin(np.clip(m[..., 2], -1, 1))
    uz, vm = ufun(lon, lat)                       # (..., nj, ni) zonal, meridional
    elon = np.stack([-np.sin(lon), np.cos(lon), np.zeros_like(lon)], -1)
    elat = np.stack([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)], -1)
    return uz * np.sum(elon * d, -1) + vm * np.sum(elat * d, -1)


def _fill_scalar_halos(ex, arrs):
    ex.scalar(arrs, cs.CENTER)


def baroclinic_wave(tiles, bounds, npz, ak, bk, consts=None, perturb=True, w_amp=0.1):
    """Returns a list of 6 dicts of native-extent state arrays (C-order (nk, nj, ni))."""
    cst = dict(CONSTANTS)
    if consts:
        cst.update(consts)
    a, omega, grav, rdgas, kappa = cst["radius"], cst["omega"], cst["grav"], cst["rdgas"], cst["kappa"]
    n, ng = bounds["ie"], bounds["ng"]
    ex = cs.Exchanger(n, ng)
    p0, u0, eta0, eta_t, T0, gamma, dT = 1.0e5, 35.0, 0.252, 0.2, 288.0, 0.005, 4.8e5
    lonc, latc, up, Rp = np.pi / 9.0, 2.0 * np.pi / 9.0, 1.0, a / 10.0
    ps = p0
    pe = ak + bk * ps                                   # (npz+1)
    peln = np.log(pe)
    delp_k = np.diff(pe)
    pm = delp_k / np.diff(peln)                         # layer-mean pressure (nh_utils.F90:440)
    eta = pm / p0
    etav = (eta - eta0) * np.pi / 2.0

    def Tbar(e):
        t = T0 * e ** (rdgas * gamma / grav)
        return np.where(e < eta_t, t + dT * np.maximum(eta_t - e, 0.0) ** 5, t)

    def T_of(lat):
        A = (-2.0 * np.sin(lat) ** 6 * (np.cos(lat) ** 2 + 1.0 / 3.0) + 10.0 / 63.0)
        B = (1.6 * np.cos(lat) ** 3 * (np.sin(lat) ** 2 + 2.0 / 3.0) - np.pi / 4.0)
        ev = etav[:, None, None]
        e = eta[:, None, None]
        return Tbar(e) + 0.75 * e * np.pi * u0 / rdgas * np.sin(ev) * np.sqrt(np.cos(ev)) * \
            (A[None] * 2.0 * u0 * np.cos(ev) ** 1.5 + B[None] * a * omega)

    def phis_of(lat):
        c = np.cos((1.0 - eta0) * np.pi / 2.0) ** 1.5
        A = (-2.0 * np.sin(lat) ** 6 * (np.cos(lat) ** 2 + 1.0 / 3.0) + 10.0 / 63.0)
        B = (1.6 * np.cos(lat) ** 3 * (np.sin(lat) ** 2 + 2.0 / 3.0) - np.pi / 4.0)
        return u0 * c * (A * u0 * c + B * a * omega)

    def wind(lon, lat):
        uz = u0 * np.cos(etav[:, None, None]) ** 1.5 * np.sin(2.0 * lat[None]) ** 2
        if perturb:
            r = a * np.arccos(np.clip(np.sin(latc) * np.sin(lat) + np.cos(latc) * np.cos(lat) * np.cos(lon - lonc), -1, 1))
            uz = uz + up * np.exp(-(r / Rp) ** 2)[None]
        return uz, np.zeros_like(uz)

    states = []
    isd, jsd = bounds["isd"], bounds["jsd"]
    nja = bounds["jed"] - jsd + 1
    nia = bounds["ied"] - isd + 1
    for g in tiles:
        lon, lat = g.arr["agrid"]
        st = {}
        T = T_of(lat)
        st["delp"] = np.broadcast_to(delp_k[:, None, None], (npz, nja, nia)).copy()
        st["pt"] = T / (pm[:, None, None] ** kappa)
        st["phis"] = phis_of(lat)[None].copy()
        kk = (np.arange(npz) + 1.0)[:, None, None]
        st["w"] = w_amp * np.sin(7.0 * lon)[None] * np.cos(5.0 * lat)[None] * np.sin(np.pi * kk / npz)
        delz = -(rdgas / grav) * T * np.diff(peln)[:, None, None]
        st["delz"] = np.ascontiguousarray(delz[:, ng:-ng, ng:-ng])
        mu, du, mv, dv = _edge_dirs(g)
        st["u"] = _wind_on_edge(mu, du, wind)
        st["v"] = _wind_on_edge(mv, dv, wind)
        states.append(st)
    # halos by exchange (overwrites the analytic halo values with the neighbours' compute-domain values)
    for nm in ("delp", "pt", "w", "phis"):
        _fill_scalar_halos(ex, [s[nm] for s in states])
    ex.pair([s["u"] for s in states], [s["v"] for s in states], cs.NORTH, cs.EAST, kind="vector")
    for s in states:
        for nm in ("delp", "pt"):
            _patch_corners(s[nm], ng, n)
    return states


def _patch_corners(a, ng, n):
    """Give the (unused) ng x ng corner blocks benign finite values (nearest interior corner cell)."""
    a[..., :ng, :ng] = a[..., ng:ng + 1, ng:ng + 1]
    a[..., :ng, -ng:] = a[..., ng:ng + 1, -ng - 1:-ng]
    a[..., -ng:, :ng] = a[..., -ng - 1:-ng, ng:ng + 1]
    a[..., -ng:, -ng:] = a[..., -ng - 1:-ng, -ng - 1:-ng]


def smooth_state(tiles, bounds, npz, seed=20241117, consts=None):
    """Analytic smooth-plus-step fields exercising every limiter branch (operator parity tests)."""
    cst = dict(CONSTANTS)
    if consts:
        cst.update(consts)
    n, ng = bounds["ie"], bounds["ng"]
    ex = cs.Exchanger(n, ng)
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, size=(8,))
    states = []
    kk = (np.arange(npz) + 1.0)[:, None, None]

    def wind(lon, lat):
        uz = 30.0 * np.cos(lat)[None] * (1.0 + 0.3 * np.sin(3 * lon + ph[0])[None] * np.cos(kk)) + 0 * kk
        vm = 12.0 * np.sin(2 * lon + ph[1])[None] * np.cos(lat)[None] ** 2 * np.sin(0.5 * kk + ph[2])
        return uz, vm

    for g in tiles:
        lon, lat = g.arr["agrid"]
        st = {}
        base = 1000.0 + 300.0 * np.sin(0.37 * kk)
        st["delp"] = base * (1.0 + 0.05 * np.sin(2 * lon + ph[3])[None] * np.cos(3 * lat)[None]) \
            + 40.0 * (np.sin(5 * lon)[None] * np.cos(4 * lat + 0.1 * kk) > 0.3)
        st["pt"] = 300.0 + 20.0 * np.cos(lat)[None] * np.sin(lon + ph[4] + 0.2 * kk) + 5.0 * (lat[None] > 0.3 + 0 * kk)
        st["w"] = 0.5 * np.sin(4 * lon + ph[5])[None] * np.cos(3 * lat)[None] * np.sin(np.pi * kk / npz)
        st["q_con"] = 1e-3 * (1.0 + np.sin(3 * lon + ph[6])[None] * np.cos(2 * lat)[None]) + 0 * kk
        mu, du, mv, dv = _edge_dirs(g)
        st["u"] = _wind_on_edge(mu, du, wind)
        st["v"] = _wind_on_edge(mv, dv, wind)
        states.append(st)
    for nm in ("delp", "pt", "w", "q_con"):
        _fill_scalar_halos(ex, [s[nm] for s in states])
    ex.pair([s["u"] for s in states], [s["v"] for s in states], cs.NORTH, cs.EAST, kind="vector")
    for s in states:
        for nm in ("delp", "pt", "w", "q_con"):
            _patch_corners(s[nm], ng, n)
    return states


# ---- SW_DYNAMICS test case 1: cosine bell in solid-body rotation (BASELINE config 1a) ----------------------
def great_circle_dist(lon1, lat1, lon2, lat2, radius):
    """fv_grid_utils.F90:1974-1996 (haversine form)."""
    beta = 2.0 * np.arcsin(np.sqrt(np.sin((lat1 - lat2) / 2.0) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon1 - lon2) / 2.0) ** 2))
    return radius * beta


def cosine_bell_height(lon, lat, radius, lon_c=0.5 * np.pi, lat_c=0.0, gh0=1.0):
    """test_cases.F90:923-941 (case 1): h = gh0/2 (1 + cos(pi r / r0)) inside r0 = radius/3 of the centre, 0 outside."""
    r0 = radius / 3.0
    r = great_circle_dist(lon_c, lat_c, lon, lat, radius)
    return np.where(r < r0, gh0 * 0.5 * (1.0 + np.cos(np.pi * r / r0)), 0.0)


def cosine_bell(tiles, bounds, alpha=0.0, consts=None):
    """Williamson test case 1 as the SW_DYNAMICS build initialises it (test_cases.F90:923-942 + init_winds with
    defOnGrid = 1, :211-503): delp = bell height (npz = 1), C-grid winds uc, vc from differences of the stream function
    psi_b = -Ubar R (sin(lat) cos(alpha) - cos(lon) cos(lat) sin(alpha)) at the cell corners (:330-335, :405-420), halo by the
    C-grid vector exchange (:421).  Ubar = 2 pi R / 12 days (:924).  Returns 6 dicts of native-extent arrays (1, nj, ni);
    u, v are the D-grid analogue from psi at the cell centres (defOnGrid = 2 formulas, :428-443) for completeness, pt = 1."""
    cst = dict(CONSTANTS)
    if consts:
        cst.update(consts)
    R = cst["radius"]
    ubar = 2.0 * np.pi * R / (12.0 * 86400.0)
    n, ng = bounds["ie"], bounds["ng"]
    ex = cs.Exchanger(n, ng)

    def psi_of(lon, lat):
        return -ubar * R * (np.sin(lat) * np.cos(alpha) - np.cos(lon) * np.cos(lat) * np.sin(alpha))

    states = []
    for g in tiles:
        lon, lat = g.arr["agrid"]
        lonb, latb = g.arr["grid"]
        st = {}
        st["delp"] = cosine_bell_height(lon, lat, R)[None].copy()
        st["pt"] = np.ones_like(st["delp"])
        psi_b = psi_of(lonb, latb)                                    # (nj+1, ni+1)
        psi = psi_of(lon, lat)                                        # (nj, ni)
        dx, dy, dxc, dyc = g.arr["dx"], g.arr["dy"], g.arr["dxc"], g.arr["dyc"]
        with np.errstate(divide="ignore", invalid="ignore"):
            vc = (psi_b[:, 1:] - psi_b[:, :-1]) / dx                  # (nj+1, ni)   :405-411
            uc = -(psi_b[1:, :] - psi_b[:-1, :]) / dy                 # (nj, ni+1)   :412-418
            v = np.zeros_like(uc); u = np.zeros_like(vc)
            v[:, 1:-1] = (psi[:, 1:] - psi[:, :-1]) / dxc[:, 1:-1]    # :428-434
            u[1:-1, :] = -(psi[1:, :] - psi[:-1, :]) / dyc[1:-1, :]   # :435-441
        for a in (uc, vc, u, v):
            a[~np.isfinite(a)] = 0.0
        st["uc_direct"], st["vc_direct"] = uc[None].copy(), vc[None].copy()   # analytic everywhere incl. the halo (tests)
        # the reference computes the compute domain only and fills the halo by the exchange
        ucc, vcc = np.zeros_like(uc), np.zeros_like(vc)
        ucc[ng:ng + n, ng:ng + n + 1] = uc[ng:ng + n, ng:ng + n + 1]
        vcc[ng:ng + n + 1, ng:ng + n] = vc[ng:ng + n + 1, ng:ng + n]
        st["uc"], st["vc"] = ucc[None].copy(), vcc[None].copy()
        st["u"], st["v"] = u[None].copy(), v[None].copy()
        states.append(st)
    ex.scalar([s["delp"] for s in states], cs.CENTER)
    ex.scalar([s["pt"] for s in states], cs.CENTER)
    ex.pair([s["uc"] for s in states], [s["vc"] for s in states], cs.EAST, cs.NORTH, kind="vector")
    ex.pair([s["u"] for s in states], [s["v"] for s in states], cs.NORTH, cs.EAST, kind="vector")
    for s in states:
        _patch_corners(s["delp"], ng, n)
    return states
