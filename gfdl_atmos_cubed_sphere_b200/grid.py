"""Equidistant gnomonic cubed-sphere grid + metric terms (host-side fixture generator).

Restates, in NumPy, the one-time initialisation the hot path depends on:

* ``model/fv_grid_utils.F90:1233-1352`` ``gnomonic_grids`` / ``gnomonic_ed`` (tile-1 corners),
* ``tools/fv_grid_tools.F90:640-693`` tile rotation (``mirror_grid :2625``) and ``shift_fac``,
* ``tools/fv_grid_tools.F90:725-1015`` ``init_grid``: halo of ``grid``, ``dx,dy,dxa,dya,dxc,dyc``,
  ``agrid``, ``area`` (``grid_area :2397``), ``area_c`` incl. the face-edge overrides ``:873-934``,
* ``model/fv_grid_utils.F90:84-790`` ``grid_utils_init``: ``sin_sg/cos_sg`` (9-point super grid),
  ``cosa*/sina*/rsin*``, edge overrides ``:533-561``, corner patches ``:373-401,577-612``,
  ``divg_u/v``, ``del6_u/v`` ``:649-675``, ``edge_factors :1121-1231``, ``da_min`` ``:680-683``,
* ``tools/test_cases.F90:763-776`` Coriolis ``f0``/``fC``,
* ``tools/fv_mp_mod.F90:1024-1449`` the ``fill_corners`` family used during initialisation.

The 6 tile orientations are not transcribed from ``rot_3d``; they are *solved for*: each
tile is the rotation of tile 1 (out of the 24 cube rotations) that makes every contact of
``fv_mp_mod.F90:498-546`` coincide point-by-point with the tiles already placed.  This is never
on the GPU path: metrics are immutable after init (``fv_arrays.F90:72-74``).

Arrays are C-ordered ``(nj, ni)`` planes == Fortran ``(i, j)`` column-major, with the native
extents of ``fv_arrays.F90:1749-1878`` (so they can be handed to the C ABI as they are).
"""
from __future__ import annotations

import itertools
import numpy as np

from . import cubed_sphere as cs

BIG = 1.0e30     # fv_grid_utils.F90 big_number (64-bit build)
TINY = 1.0e-30   # tiny_number

# FMS constants_mod, GFS/SHiELD set (not in the reference repo) -- run-time params
CONSTANTS = dict(radius=6.3712e6, omega=7.2921e-5, grav=9.80665, rdgas=287.05, cp_air=1004.6,
                 kappa=287.05 / 1004.6, pi=np.pi, rvgas=461.5)


class FA:
    """Fortran-indexed array: element (i, j) with explicit lower bounds, stored (nj, ni)."""

    def __init__(self, ilo, ihi, jlo, jhi, fill=0.0, lead=()):
        self.ilo, self.ihi, self.jlo, self.jhi = ilo, ihi, jlo, jhi
        self.a = np.full(tuple(lead) + (jhi - jlo + 1, ihi - ilo + 1), fill, dtype=np.float64)

    def s(self, i0, i1, j0, j1):
        """View of the inclusive Fortran section (i0:i1, j0:j1)."""
        return self.a[..., j0 - self.jlo:j1 - self.jlo + 1, i0 - self.ilo:i1 - self.ilo + 1]

    def __getitem__(self, ij):
        i, j = ij
        return self.a[..., j - self.jlo, i - self.ilo]

    def __setitem__(self, ij, v):
        i, j = ij
        self.a[..., j - self.jlo, i - self.ilo] = v


# ---- spherical geometry (fv_grid_utils.F90) -------------------------------------------------

def latlon2xyz(lon, lat):  # :1582
    return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)


def cart_to_latlon(p):  # :1682
    p = p / np.linalg.norm(p, axis=-1, keepdims=True)
    lon = np.where((np.abs(p[..., 0]) + np.abs(p[..., 1])) < 1e-10, 0.0, np.arctan2(p[..., 1], p[..., 0]))
    lon = np.where(lon < 0.0, 2.0 * np.pi + lon, lon)
    lat = np.arcsin(np.clip(p[..., 2], -1.0, 1.0))
    return lon, lat


def great_circle_dist(lon1, lat1, lon2, lat2, radius=1.0):  # :1974
    beta = 2.0 * np.arcsin(np.sqrt(np.sin((lat1 - lat2) / 2.0) ** 2 +
                                   np.cos(lat1) * np.cos(lat2) * np.sin((lon1 - lon2) / 2.0) ** 2))
    return radius * beta


def mid_pt3(p1, p2):  # :1930
    e = p1 + p2
    return e / np.linalg.norm(e, axis=-1, keepdims=True)


def cos_angle(p1, p2, p3):  # :2831  angle at p1 between p1->p2 and p1->p3
    P = np.cross(p1, p2)
    Q = np.cross(p1, p3)
    ddd = np.sqrt(np.sum(P * P, -1) * np.sum(Q * Q, -1))
    return np.where(ddd > 0.0, np.sum(P * Q, -1) / np.where(ddd > 0, ddd, 1.0), 1.0)


def spherical_angle(p1, p2, p3):  # :2771
    return np.arccos(np.clip(cos_angle(p1, p2, p3), -1.0, 1.0))


def quad_area(p_ll, p_ul, p_lr, p_ur, radius):
    """get_area(p1=ll, p4=ul, p2=lr, p3=ur) (:2682), points as xyz."""
    a1 = spherical_angle(p_ll, p_lr, p_ul)
    a2 = spherical_angle(p_lr, p_ur, p_ll)
    a3 = spherical_angle(p_ur, p_ul, p_lr)
    a4 = spherical_angle(p_ul, p_ur, p_ll)
    return (a1 + a2 + a3 + a4 - 2.0 * np.pi) * radius ** 2


# ---- fill_corners family (fv_mp_mod.F90:1024-1449), full-face tile so all four corners ------

def fill_corners_bgrid(q: FA, npx, npy, ng, xdir=True):  # :1031-1062
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            if xdir:
                q[1 - i, 1 - j] = q[1 - j, i + 1]
                q[1 - i, npy + j] = q[1 - j, npy - i]
                q[npx + i, 1 - j] = q[npx + j, i + 1]
                q[npx + i, npy + j] = q[npx + j, npy - i]
            else:
                q[1 - j, 1 - i] = q[i + 1, 1 - j]
                q[1 - j, npy + i] = q[i + 1, npy + j]
                q[npx + j, 1 - i] = q[npx - i, 1 - j]
                q[npx + j, npy + i] = q[npx - i, npy + j]


def fill_corners_agrid_scalar(q: FA, npx, npy, ng, xdir=True):  # :1063-1094
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            if xdir:
                q[1 - i, 1 - j] = q[1 - j, i]
                q[1 - i, npy - 1 + j] = q[1 - j, npy - 1 - i + 1]
                q[npx - 1 + i, 1 - j] = q[npx - 1 + j, i]
                q[npx - 1 + i, npy - 1 + j] = q[npx - 1 + j, npy - 1 - i + 1]
            else:
                q[1 - j, 1 - i] = q[i, 1 - j]
                q[1 - j, npy - 1 + i] = q[i, npy - 1 + j]
                q[npx - 1 + j, 1 - i] = q[npx - 1 - i + 1, 1 - j]
                q[npx - 1 + j, npy - 1 + i] = q[npx - 1 - i + 1, npy - 1 + j]


def fill_corners_dgrid(x: FA, y: FA, npx, npy, ng, sgn=1.0):  # :1249-1281
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            x[1 - i, 1 - j] = sgn * y[1 - j, i]
            x[1 - i, npy + j] = y[1 - j, npy - i]
            x[npx - 1 + i, 1 - j] = y[npx + j, i]
            x[npx - 1 + i, npy + j] = sgn * y[npx + j, npy - i]
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            y[1 - i, 1 - j] = sgn * x[j, 1 - i]
            y[1 - i, npy - 1 + j] = x[j, npy + i]
            y[npx + i, 1 - j] = x[npx - j, 1 - i]
            y[npx + i, npy - 1 + j] = sgn * x[npx - j, npy + i]


def fill_corners_cgrid(x: FA, y: FA, npx, npy, ng, sgn=1.0):  # :1361-1385
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            x[1 - i, 1 - j] = y[j, 1 - i]
            x[1 - i, npy - 1 + j] = sgn * y[j, npy + i]
            x[npx + i, 1 - j] = sgn * y[npx - j, 1 - i]
            x[npx + i, npy - 1 + j] = y[npx - j, npy + i]
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            y[1 - i, 1 - j] = x[1 - j, i]
            y[1 - i, npy + j] = sgn * x[1 - j, npy - i]
            y[npx - 1 + i, 1 - j] = sgn * x[npx + j, i]
            y[npx - 1 + i, npy + j] = x[npx + j, npy - i]


def fill_corners_agrid_pair(x: FA, y: FA, npx, npy, ng, sgn=1.0):  # :1425-1449
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            x[1 - i, 1 - j] = sgn * y[1 - j, i]
            x[1 - i, npy - 1 + j] = y[1 - j, npy - 1 - i + 1]
            x[npx - 1 + i, 1 - j] = y[npx - 1 + j, i]
            x[npx - 1 + i, npy - 1 + j] = sgn * y[npx - 1 + j, npy - 1 - i + 1]
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            y[1 - j, 1 - i] = sgn * x[i, 1 - j]
            y[1 - j, npy - 1 + i] = x[i, npy - 1 + j]
            y[npx - 1 + j, 1 - i] = x[npx - 1 - i + 1, 1 - j]
            y[npx - 1 + j, npy - 1 + i] = sgn * x[npx - 1 - i + 1, npy - 1 + j]


def fill_ghost(q: FA, npx, npy, value):  # fv_grid_utils.F90:3043
    ng = 1 - q.ilo
    q.s(q.ilo, 0, q.jlo, 0)[...] = value
    q.s(npx, npx - 1 + ng, q.jlo, 0)[...] = value
    q.s(npx, npx - 1 + ng, npy, npy - 1 + ng)[...] = value
    q.s(q.ilo, 0, npy, npy - 1 + ng)[...] = value


# ---- tile placement --------------------------------------------------------------------------

def _tile1_xyz(n):
    """gnomonic_ed (fv_grid_utils.F90:1256-1352) after symm_ed and lon -= pi (:1244-1250)."""
    rsq3 = 1.0 / np.sqrt(3.0)
    alpha = np.arcsin(rsq3)
    theta = -alpha + (2.0 * alpha / n) * np.arange(n + 1)
    t = rsq3 * np.sqrt(2.0) * np.tan(theta)
    t = 0.5 * (t - t[::-1])      # symm_ed: exact antisymmetry
    Y, Z = np.meshgrid(t, t)     # [j, i]: Y varies with i, Z with j
    P = np.stack([np.full_like(Y, rsq3), Y, Z], axis=-1)
    return P / np.linalg.norm(P, axis=-1, keepdims=True)


def _rotations():
    rots = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1.0, -1.0), repeat=3):
            R = np.zeros((3, 3))
            for r in range(3):
                R[r, perm[r]] = signs[r]
            if np.linalg.det(R) > 0:
                rots.append(R)
    return rots


def _edge_pts(P, edge):
    if edge == cs.W:
        return P[:, 0]
    if edge == cs.E:
        return P[:, -1]
    if edge == cs.S:
        return P[0, :]
    return P[-1, :]


def tile_corner_xyz(n, shift_fac=18.0):
    """6 arrays (n+1, n+1, 3): unit vectors of the cell corners of every tile."""
    P1 = _tile1_xyz(n)
    tiles = {1: P1}
    rots = _rotations()
    for t in range(2, 7):
        found = None
        for R in rots:
            cand = P1 @ R.T
            ok = True
            used = False
            for a, ea, b, eb, rev in cs.CONTACTS:
                if t == b and a in tiles:
                    ref, mine_e, ref_e = tiles[a], eb, ea
                elif t == a and b in tiles:
                    ref, mine_e, ref_e = tiles[b], ea, eb
                else:
                    continue
                used = True
                pa = _edge_pts(ref, ref_e)
                pb = _edge_pts(cand, mine_e)
                if rev:
                    pb = pb[::-1]
                if not np.allclose(pa, pb, atol=1e-12):
                    ok = False
                    break
            if ok and used:
                assert found is None, "tile orientation not unique"
                found = cand
        assert found is not None, f"no rotation places tile {t}"
        tiles[t] = found
    out = []
    ang = -np.pi / shift_fac if shift_fac > 1e-4 else 0.0   # fv_grid_tools.F90:660-661
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]])
    for t in range(1, 7):
        out.append(tiles[t] @ Rz.T)
    # make shared edges bitwise identical (fv_grid_tools.F90:677-693): copy from the lower tile
    for a, ea, b, eb, rev in cs.CONTACTS:
        pa = _edge_pts(out[a - 1], ea)
        src = pa[::-1] if rev else pa
        _edge_pts(out[b - 1], eb)[...] = src
    return out


# ---- the generator ---------------------------------------------------------------------------

class TileGrid:
    """All metric arrays of one tile, native extents (dict name -> ndarray) + scalars."""

    def __init__(self):
        self.arr = {}
        self.da_min = 0.0
        self.da_min_c = 0.0


def make_cubed_sphere(n, ng=3, consts=None, alpha=0.0, shift_fac=18.0):
    """Return (list of 6 TileGrid, bounds dict). n = npx-1 cells per edge."""
    cst = dict(CONSTANTS)
    if consts:
        cst.update(consts)
    radius, omega = cst["radius"], cst["omega"]
    npx = npy = n + 1
    is_, ie, js, je = 1, n, 1, n
    isd, ied, jsd, jed = 1 - ng, n + ng, 1 - ng, n + ng
    ex = cs.Exchanger(n, ng)
    corners = tile_corner_xyz(n, shift_fac)

    # --- grid (corner xyz with halo): CORNER exchange + fill_corners(BGRID, XDir)  fv_grid_tools.F90:725-729
    G3 = [FA(isd, ied + 1, jsd, jed + 1, lead=(3,)) for _ in range(6)]
    for t in range(6):
        G3[t].s(1, npx, 1, npy)[...] = np.moveaxis(corners[t], -1, 0)
    ex.scalar([g.a for g in G3], cs.CORNER)
    for t in range(6):
        fill_corners_bgrid(G3[t], npx, npy, ng, xdir=True)
    grid_xyz = [np.moveaxis(g.a, 0, -1).copy() for g in G3]          # (nj+1, ni+1, 3)
    glon, glat = [], []
    for t in range(6):
        lo, la = cart_to_latlon(np.where(np.abs(grid_xyz[t]).sum(-1, keepdims=True) > 0, grid_xyz[t], [1.0, 0, 0]))
        glon.append(lo)
        glat.append(la)

    def gsec(t, i0, i1, j0, j1):
        return grid_xyz[t][j0 - jsd:j1 - jsd + 1, i0 - isd:i1 - isd + 1]

    def gll(t, i0, i1, j0, j1):
        return (glon[t][j0 - jsd:j1 - jsd + 1, i0 - isd:i1 - isd + 1], glat[t][j0 - jsd:j1 - jsd + 1, i0 - isd:i1 - isd + 1])

    def gcd_xyz(p, q):
        lo1, la1 = cart_to_latlon(p)
        lo2, la2 = cart_to_latlon(q)
        return great_circle_dist(lo1, la1, lo2, la2, radius)

    # --- dx, dy on the compute domain, SCALAR_PAIR halo, fill_corners(DGRID)  :744-783
    dx = [FA(isd, ied, jsd, jed + 1) for _ in range(6)]
    dy = [FA(isd, ied + 1, jsd, jed) for _ in range(6)]
    for t in range(6):
        lo, la = gll(t, is_, ie + 1, js, je + 1)
        dx[t].s(is_, ie, js, je + 1)[...] = great_circle_dist(lo[:, 1:], la[:, 1:], lo[:, :-1], la[:, :-1], radius)
        dy[t].s(is_, ie + 1, js, je)[...] = great_circle_dist(lo[1:, :], la[1:, :], lo[:-1, :], la[:-1, :], radius)
    ex.pair([a.a for a in dy], [a.a for a in dx], cs.EAST, cs.NORTH, kind="pair")
    for t in range(6):
        fill_corners_dgrid(dx[t], dy[t], npx, npy, ng, 1.0)

    # --- agrid: cell centres + CENTER halo + fill_corners  :790-812
    alon = [FA(isd, ied, jsd, jed, fill=-1e25) for _ in range(6)]
    alat = [FA(isd, ied, jsd, jed, fill=-1e25) for _ in range(6)]
    for t in range(6):
        c = gsec(t, is_, ie + 1, js, je + 1)
        ec = c[:-1, :-1] + c[:-1, 1:] + c[1:, :-1] + c[1:, 1:]
        lo, la = cart_to_latlon(ec)
        alon[t].s(is_, ie, js, je)[...] = lo
        alat[t].s(is_, ie, js, je)[...] = la
    ex.scalar([a.a for a in alon], cs.CENTER)
    ex.scalar([a.a for a in alat], cs.CENTER)
    for t in range(6):
        fill_corners_agrid_scalar(alon[t], npx, npy, ng, xdir=True)
        fill_corners_agrid_scalar(alat[t], npx, npy, ng, xdir=False)

    # --- dxa, dya over the data domain + fill_corners(AGRID)  :814-828
    dxa = [FA(isd, ied, jsd, jed) for _ in range(6)]
    dya = [FA(isd, ied, jsd, jed) for _ in range(6)]
    for t in range(6):
        c = gsec(t, isd, ied + 1, jsd, jed + 1)
        with np.errstate(invalid="ignore", divide="ignore"):
            p1 = mid_pt3(c[:-1, :-1], c[1:, :-1])
            p2 = mid_pt3(c[:-1, 1:], c[1:, 1:])
            dxa[t].a[...] = gcd_xyz(p2, p1)
            p1 = mid_pt3(c[:-1, :-1], c[:-1, 1:])
            p2 = mid_pt3(c[1:, :-1], c[1:, 1:])
            dya[t].a[...] = gcd_xyz(p2, p1)
        fill_corners_agrid_pair(dxa[t], dya[t], npx, npy, ng, 1.0)

    # --- dxc, dyc  :836-943
    dxc = [FA(isd, ied + 1, jsd, jed) for _ in range(6)]
    dyc = [FA(isd, ied, jsd, jed + 1) for _ in range(6)]
    axyz = [latlon2xyz(alon[t].a, alat[t].a) for t in range(6)]
    for t in range(6):
        lo, la = alon[t].a, alat[t].a
        dxc[t].s(isd + 1, ied, jsd, jed)[...] = great_circle_dist(lo[:, 1:], la[:, 1:], lo[:, :-1], la[:, :-1], radius)
        dxc[t].s(isd, isd, jsd, jed)[...] = dxc[t].s(isd + 1, isd + 1, jsd, jed)
        dxc[t].s(ied + 1, ied + 1, jsd, jed)[...] = dxc[t].s(ied, ied, jsd, jed)
        dyc[t].s(isd, ied, jsd + 1, jed)[...] = great_circle_dist(lo[1:, :], la[1:, :], lo[:-1, :], la[:-1, :], radius)
        dyc[t].s(isd, ied, jsd, jsd)[...] = dyc[t].s(isd, ied, jsd + 1, jsd + 1)
        dyc[t].s(isd, ied, jed + 1, jed + 1)[...] = dyc[t].s(isd, ied, jed, jed)

    # --- area (:2440-2455) and area_c (:2475-2493 + corner triangles, then edge overrides :873-934)
    area = [FA(isd, ied, jsd, jed) for _ in range(6)]
    area_c = [FA(isd, ied + 1, jsd, jed + 1) for _ in range(6)]

    def A(t, i0, i1, j0, j1):
        return axyz[t][j0 - jsd:j1 - jsd + 1, i0 - isd:i1 - isd + 1]

    for t in range(6):
        c = gsec(t, is_, ie + 1, js, je + 1)
        area[t].s(is_, ie, js, je)[...] = quad_area(c[:-1, :-1], c[1:, :-1], c[:-1, 1:], c[1:, 1:], radius)
        a = A(t, is_ - 1, ie + 1, js - 1, je + 1)
        with np.errstate(invalid="ignore"):
            area_c[t].s(is_, ie + 1, js, je + 1)[...] = quad_area(a[:-1, :-1], a[1:, :-1], a[:-1, 1:], a[1:, 1:], radius)
        # edge overrides, in the reference's order (the four cube-vertex values end up as the
        # last writer's 2 x half-cell value, fv_grid_tools.F90:875-934)
        g = gsec
        i = 1
        p1 = mid_pt3(g(t, i, i, js - 1, je)[:, 0], g(t, i, i, js, je + 1)[:, 0])
        p4 = mid_pt3(g(t, i, i, js, je + 1)[:, 0], g(t, i, i, js + 1, je + 2)[:, 0])
        p2 = A(t, i, i, js - 1, je)[:, 0]
        p3 = A(t, i, i, js, je + 1)[:, 0]
        area_c[t].s(i, i, js, je + 1)[:, 0] = 2.0 * quad_area(p1, p4, p2, p3, radius)
        pm = mid_pt3(g(t, i, i, js, je)[:, 0], g(t, i, i, js + 1, je + 1)[:, 0])
        dxc[t].s(i, i, js, je)[:, 0] = 2.0 * gcd_xyz(pm, A(t, i, i, js, je)[:, 0])
        i = npx
        p1 = A(t, i - 1, i - 1, js - 1, je)[:, 0]
        p2 = mid_pt3(g(t, i, i, js - 1, je)[:, 0], g(t, i, i, js, je + 1)[:, 0])
        p3 = mid_pt3(g(t, i, i, js, je + 1)[:, 0], g(t, i, i, js + 1, je + 2)[:, 0])
        p4 = A(t, i - 1, i - 1, js, je + 1)[:, 0]
        area_c[t].s(i, i, js, je + 1)[:, 0] = 2.0 * quad_area(p1, p4, p2, p3, radius)
        pm = mid_pt3(g(t, i, i, js, je)[:, 0], g(t, i, i, js + 1, je + 1)[:, 0])
        dxc[t].s(i, i, js, je)[:, 0] = 2.0 * gcd_xyz(A(t, i - 1, i - 1, js, je)[:, 0], pm)
        j = 1
        p1 = mid_pt3(g(t, is_ - 1, ie, j, j)[0], g(t, is_, ie + 1, j, j)[0])
        p2 = mid_pt3(g(t, is_, ie + 1, j, j)[0], g(t, is_ + 1, ie + 2, j, j)[0])
        p3 = A(t, is_, ie + 1, j, j)[0]
        p4 = A(t, is_ - 1, ie, j, j)[0]
        area_c[t].s(is_, ie + 1, j, j)[0] = 2.0 * quad_area(p1, p4, p2, p3, radius)
        pm = mid_pt3(g(t, is_, ie, j, j)[0], g(t, is_ + 1, ie + 1, j, j)[0])
        dyc[t].s(is_, ie, j, j)[0] = 2.0 * gcd_xyz(pm, A(t, is_, ie, j, j)[0])
        j = npy
        p1 = A(t, is_ - 1, ie, j - 1, j - 1)[0]
        p2 = A(t, is_, ie + 1, j - 1, j - 1)[0]
        p3 = mid_pt3(g(t, is_, ie + 1, j, j)[0], g(t, is_ + 1, ie + 2, j, j)[0])
        p4 = mid_pt3(g(t, is_ - 1, ie, j, j)[0], g(t, is_, ie + 1, j, j)[0])
        area_c[t].s(is_, ie + 1, j, j)[0] = 2.0 * quad_area(p1, p4, p2, p3, radius)
        pm = mid_pt3(g(t, is_, ie, j, j)[0], g(t, is_ + 1, ie + 1, j, j)[0])
        dyc[t].s(is_, ie, j, j)[0] = 2.0 * gcd_xyz(A(t, is_, ie, j - 1, j - 1)[0], pm)
    ex.pair([a.a for a in dxc], [a.a for a in dyc], cs.EAST, cs.NORTH, kind="pair")   # :939
    for t in range(6):
        fill_corners_cgrid(dxc[t], dyc[t], npx, npy, ng, 1.0)                           # :942
    ex.scalar([a.a for a in area], cs.CENTER)                                          # :945
    ex.scalar([a.a for a in area_c], cs.CORNER)                                        # :976
    for t in range(6):
        fill_ghost(area[t], npx, npy, -BIG)                                            # :980
        fill_corners_bgrid(area_c[t], npx, npy, ng, xdir=True)                         # :981

    # ================= grid_utils_init (fv_grid_utils.F90:84-790) =================
    tiles = []
    for t in range(6):
        T = TileGrid()
        g3 = grid_xyz[t]                                    # (isd:ied+1, jsd:jed+1)
        c00, c10, c01, c11 = g3[:-1, :-1], g3[:-1, 1:], g3[1:, :-1], g3[1:, 1:]   # (i,j),(i+1,j),(i,j+1),(i+1,j+1)
        cos_sg = FA(isd, ied, jsd, jed, fill=BIG, lead=(9,))
        sin_sg = FA(isd, ied, jsd, jed, fill=TINY, lead=(9,))
        with np.errstate(invalid="ignore", divide="ignore"):
            # get_center_vect :1738
            pc = c00 + c10 + c01 + c11
            pc = pc / np.linalg.norm(pc, axis=-1, keepdims=True)
            p3v = np.cross(mid_pt3(c10, c11), mid_pt3(c00, c01))
            ec1 = np.cross(pc, p3v)
            ec1 = ec1 / np.linalg.norm(ec1, axis=-1, keepdims=True)
            p3v = np.cross(mid_pt3(c01, c11), mid_pt3(c00, c10))
            ec2 = np.cross(pc, p3v)
            ec2 = ec2 / np.linalg.norm(ec2, axis=-1, keepdims=True)
            p3 = axyz[t]
            cs_ = cos_sg.a
            cs_[5] = cos_angle(c00, c10, c01)              # 6: SW corner
            cs_[6] = -cos_angle(c10, c00, c11)             # 7: SE
            cs_[7] = cos_angle(c11, c10, c01)              # 8: NE
            cs_[8] = -cos_angle(c01, c00, c11)             # 9: NW
            cs_[0] = cos_angle(mid_pt3(c00, c01), p3, c01)  # 1: W mid
            cs_[1] = cos_angle(mid_pt3(c00, c10), c10, p3)  # 2: S mid
            cs_[2] = cos_angle(mid_pt3(c10, c11), p3, c10)  # 3: E mid
            cs_[3] = cos_angle(mid_pt3(c01, c11), c01, p3)  # 4: N mid
            cs_[4] = np.sum(ec1 * ec2, axis=-1)             # 5: centre
        cs_[~np.isfinite(cs_)] = BIG
        with np.errstate(over="ignore", invalid="ignore"):
            sin_sg.a[...] = np.minimum(1.0, np.sqrt(np.maximum(0.0, 1.0 - cs_ ** 2)))

        def SS(i, j, k):
            return sin_sg.a[k - 1, j - jsd, i - isd]

        def setS(i, j, k, v):
            sin_sg.a[k - 1, j - jsd, i - isd] = v

        def CC(i, j, k):
            return cos_sg.a[k - 1, j - jsd, i - isd]

        def setC(i, j, k, v):
            cos_sg.a[k - 1, j - jsd, i - isd] = v

        # first corner patch (sin only) :373-401
        for i in range(-2, 1):
            setS(0, i, 3, SS(i, 1, 2)); setS(i, 0, 4, SS(1, i, 1))
        for i in range(npy, npy + 3):
            setS(0, i, 3, SS(npy - i, npy - 1, 4))
        for i in range(-2, 1):
            setS(i, npy, 2, SS(1, npx + i, 1))
        for j in range(-2, 1):
            setS(npx, j, 1, SS(npx - j, 1, 2))
        for i in range(npx, npx + 3):
            setS(i, 0, 4, SS(npx - 1, npx - i, 3))
        for i in range(npy, npy + 3):
            setS(npx, i, 1, SS(i, npy - 1, 4)); setS(i, npy, 2, SS(npx - 1, i, 3))

        cosa = FA(isd, ied + 1, jsd, jed + 1, fill=BIG)
        sina = FA(isd, ied + 1, jsd, jed + 1, fill=BIG)
        cosa.s(is_, ie + 1, js, je + 1)[...] = 0.5 * (cos_sg.a[7][js - 1 - jsd:je + 1 - jsd, is_ - 1 - isd:ie + 1 - isd] +
                                                       cos_sg.a[5][js - jsd:je + 2 - jsd, is_ - isd:ie + 2 - isd])
        sina.s(is_, ie + 1, js, je + 1)[...] = 0.5 * (sin_sg.a[7][js - 1 - jsd:je + 1 - jsd, is_ - 1 - isd:ie + 1 - isd] +
                                                       sin_sg.a[5][js - jsd:je + 2 - jsd, is_ - isd:ie + 2 - isd])
        cosa_u = FA(isd, ied + 1, jsd, jed, fill=BIG); sina_u = FA(isd, ied + 1, jsd, jed, fill=BIG)
        rsin_u = FA(isd, ied + 1, jsd, jed, fill=BIG)
        cosa_v = FA(isd, ied, jsd, jed + 1, fill=BIG); sina_v = FA(isd, ied, jsd, jed + 1, fill=BIG)
        rsin_v = FA(isd, ied, jsd, jed + 1, fill=BIG)
        with np.errstate(over="ignore", invalid="ignore"):
            cosa_u.s(isd + 1, ied, jsd, jed)[...] = 0.5 * (cos_sg.a[2][:, :-1] + cos_sg.a[0][:, 1:])
            sina_u.s(isd + 1, ied, jsd, jed)[...] = 0.5 * (sin_sg.a[2][:, :-1] + sin_sg.a[0][:, 1:])
            rsin_u.s(isd + 1, ied, jsd, jed)[...] = 1.0 / np.maximum(TINY, sina_u.s(isd + 1, ied, jsd, jed) ** 2)
            cosa_v.s(isd, ied, jsd + 1, jed)[...] = 0.5 * (cos_sg.a[3][:-1, :] + cos_sg.a[1][1:, :])
            sina_v.s(isd, ied, jsd + 1, jed)[...] = 0.5 * (sin_sg.a[3][:-1, :] + sin_sg.a[1][1:, :])
            rsin_v.s(isd, ied, jsd + 1, jed)[...] = 1.0 / np.maximum(TINY, sina_v.s(isd, ied, jsd + 1, jed) ** 2)
            cosa_s = FA(isd, ied, jsd, jed); rsin2 = FA(isd, ied, jsd, jed)
            cosa_s.a[...] = cos_sg.a[4]
            rsin2.a[...] = 1.0 / np.maximum(TINY, sin_sg.a[4] ** 2)
        fill_ghost(cosa_s, npx, npy, BIG)                                                  # :528
        rsina = FA(is_, ie + 1, js, je + 1, fill=BIG)                                      # :533-543
        with np.errstate(over="ignore"):
            rsina.s(2, npx - 1, 2, npy - 1)[...] = 1.0 / np.maximum(TINY, sina.s(2, npx - 1, 2, npy - 1) ** 2)
        for i in (1, npx):                                                                 # :545-552
            v = sina_u.s(i, i, jsd, jed)
            rsin_u.s(i, i, jsd, jed)[...] = 1.0 / (np.sign(v) * np.maximum(TINY, np.abs(v)) + (v == 0) * TINY)
        for j in (1, npy):                                                                 # :554-561
            v = sina_v.s(isd, ied, j, j)
            rsin_v.s(isd, ied, j, j)[...] = 1.0 / (np.sign(v) * np.maximum(TINY, np.abs(v)) + (v == 0) * TINY)
        for k in range(9):                                                                 # :567-572
            tmpS = FA(isd, ied, jsd, jed); tmpS.a = sin_sg.a[k]; fill_ghost(tmpS, npx, npy, TINY)
            tmpC = FA(isd, ied, jsd, jed); tmpC.a = cos_sg.a[k]; fill_ghost(tmpC, npx, npy, BIG)
        # second corner patch (sin and cos) :577-612
        for i in range(0, -3, -1):
            setS(0, i, 3, SS(i, 1, 2)); setS(i, 0, 4, SS(1, i, 1))
            setC(0, i, 3, CC(i, 1, 2)); setC(i, 0, 4, CC(1, i, 1))
        for i in range(npy, npy + 3):
            setS(0, i, 3, SS(npy - i, npy - 1, 4)); setC(0, i, 3, CC(npy - i, npy - 1, 4))
        for i in range(0, -3, -1):
            setS(i, npy, 2, SS(1, npy - i, 1)); setC(i, npy, 2, CC(1, npy - i, 1))
        for j in range(0, -3, -1):
            setS(npx, j, 1, SS(npx - j, 1, 2)); setC(npx, j, 1, CC(npx - j, 1, 2))
        for i in range(npx, npx + 3):
            setS(i, 0, 4, SS(npx - 1, npx - i, 3)); setC(i, 0, 4, CC(npx - 1, npx - i, 3))
        for i in range(0, 3):
            setS(npx, npy + i, 1, SS(npx + i, npy - 1, 4)); setS(npx + i, npy, 2, SS(npx - 1, npy + i, 3))
            setC(npx, npy + i, 1, CC(npx + i, npy - 1, 4)); setC(npx + i, npy, 2, CC(npx - 1, npy + i, 3))

        # divg_u/v, del6_u/v :649-675
        divg_u = FA(isd, ied, jsd, jed + 1); del6_u = FA(isd, ied, jsd, jed + 1)
        divg_v = FA(isd, ied + 1, jsd, jed); del6_v = FA(isd, ied + 1, jsd, jed)
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            divg_u.a[...] = sina_v.a * dyc[t].a / dx[t].a
            del6_u.a[...] = sina_v.a * dx[t].a / dyc[t].a
            for j in (1, npy):
                sv = 0.5 * (sin_sg.a[1][j - jsd, :] + sin_sg.a[3][j - 1 - jsd, :])
                divg_u.s(isd, ied, j, j)[0] = sv * dyc[t].s(isd, ied, j, j)[0] / dx[t].s(isd, ied, j, j)[0]
                del6_u.s(isd, ied, j, j)[0] = sv * dx[t].s(isd, ied, j, j)[0] / dyc[t].s(isd, ied, j, j)[0]
            divg_v.a[...] = sina_u.a * dxc[t].a / dy[t].a
            del6_v.a[...] = sina_u.a * dy[t].a / dxc[t].a
            for (i, ia, ib) in ((1, 1, 0), (npx, npx, npx - 1)):
                sv = 0.5 * (sin_sg.a[0][:, ia - isd] + sin_sg.a[2][:, ib - isd])
                divg_v.s(i, i, jsd, jed)[:, 0] = sv * dxc[t].s(i, i, jsd, jed)[:, 0] / dy[t].s(i, i, jsd, jed)[:, 0]
                del6_v.s(i, i, jsd, jed)[:, 0] = sv * dy[t].s(i, i, jsd, jed)[:, 0] / dxc[t].s(i, i, jsd, jed)[:, 0]

        # edge_factors :1121-1231
        def efac(pts_lon, pts_lat, glo, gla):
            d1 = great_circle_dist(pts_lon[:-1], pts_lat[:-1], glo, gla)
            d2 = great_circle_dist(pts_lon[1:], pts_lat[1:], glo, gla)
            return d2 / (d1 + d2)

        edge_w = np.full(npy, BIG); edge_e = np.full(npy, BIG); edge_s = np.full(npx, BIG); edge_n = np.full(npx, BIG)
        ax = axyz[t]
        for (name, i) in (("w", 1), ("e", npx)):
            pm = mid_pt3(ax[1 - jsd:npy - 1 - jsd + 1, i - 1 - isd], ax[1 - jsd:npy - 1 - jsd + 1, i - isd])
            lo, la = cart_to_latlon(pm)
            glo = glon[t][2 - jsd:npy - 1 - jsd + 1, i - isd]; gla = glat[t][2 - jsd:npy - 1 - jsd + 1, i - isd]
            (edge_w if name == "w" else edge_e)[1:npy - 1] = efac(lo, la, glo, gla)
        for (name, j) in (("s", 1), ("n", npy)):
            pm = mid_pt3(ax[j - 1 - jsd, 1 - isd:npx - 1 - isd + 1], ax[j - jsd, 1 - isd:npx - 1 - isd + 1])
            lo, la = cart_to_latlon(pm)
            glo = glon[t][j - jsd, 2 - isd:npx - 1 - isd + 1]; gla = glat[t][j - jsd, 2 - isd:npx - 1 - isd + 1]
            (edge_s if name == "s" else edge_n)[1:npx - 1] = efac(lo, la, glo, gla)

        # Coriolis  tools/test_cases.F90:763-776
        fC = 2.0 * omega * (-np.cos(glon[t]) * np.cos(glat[t]) * np.sin(alpha) + np.sin(glat[t]) * np.cos(alpha))
        f0 = FA(isd, ied, jsd, jed)
        f0.a[...] = 2.0 * omega * (-np.cos(alon[t].a) * np.cos(alat[t].a) * np.sin(alpha) + np.sin(alat[t].a) * np.cos(alpha))
        T._f0 = f0

        a = T.arr
        a["area"] = area[t].a
        a["dxa"] = dxa[t].a; a["dya"] = dya[t].a
        a["cosa_s"] = cosa_s.a; a["rsin2"] = rsin2.a
        a["sin_sg"] = sin_sg.a; a["cos_sg"] = cos_sg.a
        a["dy"] = dy[t].a; a["dxc"] = dxc[t].a; a["cosa_u"] = cosa_u.a; a["sina_u"] = sina_u.a; a["rsin_u"] = rsin_u.a
        a["dx"] = dx[t].a; a["dyc"] = dyc[t].a; a["cosa_v"] = cosa_v.a; a["sina_v"] = sina_v.a; a["rsin_v"] = rsin_v.a
        a["area_c"] = area_c[t].a; a["fC"] = fC; a["cosa"] = cosa.a; a["sina"] = sina.a; a["rsina"] = rsina.a
        a["edge_w"] = edge_w; a["edge_e"] = edge_e; a["edge_s"] = edge_s; a["edge_n"] = edge_n
        a["grid"] = np.stack([glon[t], glat[t]], axis=0)
        a["agrid"] = np.stack([alon[t].a, alat[t].a], axis=0)
        # unit vectors of the omega diagnostic (adv_pe, dyn_core.F90:1529-1630), Fortran-native extents with the component fastest:
        # ec1, ec2 (3, isd:ied, jsd:jed) from get_center_vect (above; zero in the corner ghost blocks, fv_grid_utils.F90:1757-1760),
        # en1 (3, is:ie, js:je+1) / en2 (3, is:ie+1, js:je) normal to the cell edges (:629-642)
        with np.errstate(invalid="ignore"):
            e1 = np.where(np.isfinite(ec1), ec1, 0.0); e2 = np.where(np.isfinite(ec2), ec2, 0.0)
        jj, ii = np.meshgrid(np.arange(jsd, jed + 1), np.arange(isd, ied + 1), indexing="ij")
        ghost = ((ii < 1) | (ii > npx - 1)) & ((jj < 1) | (jj > npy - 1))
        e1[ghost] = 0.0; e2[ghost] = 0.0
        a["ec1"] = np.ascontiguousarray(e1); a["ec2"] = np.ascontiguousarray(e2)
        gc = g3[1 - jsd:npy - jsd + 1, 1 - isd:npx - isd + 1]              # corners (1:npx, 1:npy)
        n1 = np.cross(gc[:, :-1], gc[:, 1:])                             # (js:je+1, is:ie): grid3(i,j) x grid3(i+1,j)
        n2 = np.cross(gc[1:, :], gc[:-1, :])                             # (js:je, is:ie+1): grid3(i,j+1) x grid3(i,j)
        a["en1"] = np.ascontiguousarray(n1 / np.linalg.norm(n1, axis=-1, keepdims=True))
        a["en2"] = np.ascontiguousarray(n2 / np.linalg.norm(n2, axis=-1, keepdims=True))
        T._divg = (divg_u, divg_v, del6_u, del6_v)
        tiles.append(T)

    # halo of f0 (test_cases.F90:777-778) and SCALAR_PAIR halo of divg/del6 (fv_grid_utils.F90:695-698)
    ex.scalar([T._f0.a for T in tiles], cs.CENTER)
    for T in tiles:
        fill_corners_agrid_scalar(T._f0, npx, npy, ng, xdir=False)
        T.arr["f0"] = T._f0.a
    ex.pair([T._divg[1].a for T in tiles], [T._divg[0].a for T in tiles], cs.EAST, cs.NORTH, kind="pair")
    ex.pair([T._divg[3].a for T in tiles], [T._divg[2].a for T in tiles], cs.EAST, cs.NORTH, kind="pair")
    da_min = min(float(area[t].s(is_, ie, js, je).min()) for t in range(6))           # :680
    da_min_c = min(float(area_c[t].s(is_, ie, js, je).min()) for t in range(6))       # :683
    for T in tiles:
        a = T.arr
        a["divg_u"], a["divg_v"], a["del6_u"], a["del6_v"] = (x.a for x in T._divg)
        with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
            a["rarea"] = 1.0 / a["area"]; a["rdxa"] = 1.0 / a["dxa"]; a["rdya"] = 1.0 / a["dya"]
            a["rdy"] = 1.0 / a["dy"]; a["rdxc"] = 1.0 / a["dxc"]; a["rdx"] = 1.0 / a["dx"]; a["rdyc"] = 1.0 / a["dyc"]
            a["rarea_c"] = 1.0 / a["area_c"]
        for k in list(a.keys()):
            v = a[k]
            v[~np.isfinite(v)] = BIG
            a[k] = np.ascontiguousarray(v, dtype=np.float64)
        T.da_min, T.da_min_c = da_min, da_min_c
        del T._f0, T._divg
    bounds = dict(npx=npx, npy=npy, ng=ng, is_=is_, ie=ie, js=js, je=je, isd=isd, ied=ied, jsd=jsd, jed=jed,
                  grid_type=0)
    return tiles, bounds


def make_cartesian(n, ng=3, dx_const=1000.0, deglat=15.0, consts=None):
    """Doubly-periodic Cartesian tile (grid_type=4): constant dx=dy, orthogonal (sin=1, cos=0).

    fv_grid_tools.F90:1160-1290 setup_cartesian; fv_grid_utils.F90:425-452,614-626.
    """
    cst = dict(CONSTANTS)
    if consts:
        cst.update(consts)
    npx = npy = n + 1
    isd, ied, jsd, jed = 1 - ng, n + ng, 1 - ng, n + ng
    nia, nja = ied - isd + 1, jed - jsd + 1
    T = TileGrid()
    a = T.arr
    one_a = np.ones((nja, nia))
    for nm, shp in (("area", (nja, nia)), ("dxa", (nja, nia)), ("dya", (nja, nia)), ("dy", (nja, nia + 1)),
                    ("dxc", (nja, nia + 1)), ("dx", (nja + 1, nia)), ("dyc", (nja + 1, nia)),
                    ("area_c", (nja + 1, nia + 1))):
        val = dx_const * dx_const if nm.startswith("area") else dx_const
        a[nm] = np.full(shp, val)
    a["cosa_s"] = np.zeros((nja, nia)); a["rsin2"] = one_a.copy()
    a["sin_sg"] = np.ones((9, nja, nia)); a["cos_sg"] = np.zeros((9, nja, nia))
    a["ec1"] = np.zeros((nja, nia, 3)); a["ec1"][..., 0] = 1.0             # fv_grid_utils.F90:429-435
    a["ec2"] = np.zeros((nja, nia, 3)); a["ec2"][..., 1] = 1.0
    a["en1"] = np.zeros((n + 1, n, 3)); a["en2"] = np.zeros((n, n + 1, 3))   # not set for grid_type >= 3 (:628)
    for nm, shp in (("cosa_u", (nja, nia + 1)), ("cosa_v", (nja + 1, nia)), ("cosa", (nja + 1, nia + 1))):
        a[nm] = np.zeros(shp)
    for nm, shp in (("sina_u", (nja, nia + 1)), ("rsin_u", (nja, nia + 1)), ("sina_v", (nja + 1, nia)),
                    ("rsin_v", (nja + 1, nia)), ("sina", (nja + 1, nia + 1)), ("rsina", (n + 1, n + 1)),
                    ("divg_u", (nja + 1, nia)), ("del6_u", (nja + 1, nia)), ("divg_v", (nja, nia + 1)),
                    ("del6_v", (nja, nia + 1))):
        a[nm] = np.ones(shp)
    f = 2.0 * cst["omega"] * np.sin(np.deg2rad(deglat))
    a["f0"] = np.full((nja, nia), f); a["fC"] = np.full((nja + 1, nia + 1), f)
    for nm in ("edge_w", "edge_e", "edge_s", "edge_n"):
        a[nm] = np.full(npx, BIG)
    a["grid"] = np.zeros((2, nja + 1, nia + 1)); a["agrid"] = np.zeros((2, nja, nia))
    for nm, src in (("rarea", "area"), ("rdxa", "dxa"), ("rdya", "dya"), ("rdy", "dy"), ("rdxc", "dxc"),
                    ("rdx", "dx"), ("rdyc", "dyc"), ("rarea_c", "area_c")):
        a[nm] = 1.0 / a[src]
    T.da_min = T.da_min_c = dx_const * dx_const
    bounds = dict(npx=npx, npy=npy, ng=ng, is_=1, ie=n, js=1, je=n, isd=isd, ied=ied, jsd=jsd, jed=jed, grid_type=4)
    return [T], bounds
