"""CPU: grid generator sanity and the halo tables (NumPy builder vs the
C++ builder inside the CUDA library, and both vs analytic vector fields)."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H
from gfdl_atmos_cubed_sphere_b200 import grid as G, cubed_sphere as cs, init_state as I


def test_areas_sum_to_sphere_and_dx_ratio():
    tiles, b = H.cube_grid(24)
    R = G.CONSTANTS["radius"]
    tot = sum(T.arr["area"][3:-3, 3:-3].sum() for T in tiles)
    assert abs(tot / (4 * np.pi * R * R) - 1.0) < 1e-12
    dxs = [T.arr["dx"][3:-3, 3:-3] for T in tiles]
    ratio = max(d.max() for d in dxs) / min(d.min() for d in dxs)
    assert abs(ratio - np.sqrt(2.0)) < 2e-3          # fv_grid_utils.F90:1262
    a0 = tiles[0].arr["area"][3:-3, 3:-3]
    for T in tiles[1:]:
        assert np.allclose(np.sort(T.arr["area"][3:-3, 3:-3].ravel()), np.sort(a0.ravel()), rtol=1e-12)


def test_all_twelve_contacts_share_edges():
    P = G.tile_corner_xyz(10)
    for a, ea, b_, eb, rev in cs.CONTACTS:
        pa = G._edge_pts(P[a - 1], ea); pb = G._edge_pts(P[b_ - 1], eb)
        assert np.array_equal(pa, pb[::-1] if rev else pb)


def test_vector_halo_exchange_matches_analytic_winds():
    """D-grid (u,v) and C-grid (uc,vc) halos obtained by exchange equal the analytic wind projected
    on the local edge directions of the halo cells -> validates rotation + sign at all 12 contacts."""
    n, ng = 12, 3
    tiles, b = H.cube_grid(n)

    def wind(lon, lat):
        return (30 * np.cos(lat) * (1 + 0.3 * np.sin(3 * lon)))[None], (12 * np.sin(2 * lon) * np.cos(lat) ** 2)[None]
    us, vs = [], []
    for g in tiles:
        mu, du, mv, dv = I._edge_dirs(g)
        us.append(I._wind_on_edge(mu, du, wind)); vs.append(I._wind_on_edge(mv, dv, wind))
    ua = [u.copy() for u in us]; va = [v.copy() for v in vs]
    for t in range(6):   # poison the halos so the exchange has to fill them
        us[t][0, :ng, :] = 1e9; us[t][0, -ng:, :] = 1e9; us[t][0, :, :ng] = 1e9; us[t][0, :, -ng:] = 1e9
        vs[t][0, :ng, :] = 1e9; vs[t][0, -ng:, :] = 1e9; vs[t][0, :, :ng] = 1e9; vs[t][0, :, -ng:] = 1e9
    ex = cs.Exchanger(n, ng)
    ex.pair(us, vs, cs.NORTH, cs.EAST, kind="vector")

    def strips(shape):
        m = np.zeros(shape, bool); nj, ni = shape
        m[ng:nj - ng, :] = True; m[:, ng:ni - ng] = True
        return m
    for t in range(6):
        mu_, mv_ = strips(us[t].shape[1:]), strips(vs[t].shape[1:])
        assert np.abs((us[t] - ua[t])[0][mu_]).max() < 1e-12
        assert np.abs((vs[t] - va[t])[0][mv_]).max() < 1e-12


def _lib_tables(lib, n, ng, tile, ncomp, posx, posy, ci, vector, halo, bonly):
    cap = 8 * (n + 8) * 4 * 2
    arrs = [np.zeros(cap, dtype=np.int32) for _ in range(5)]
    ip = C.POINTER(C.c_int)
    cnt = lib.fv3_halo_entries(n + 1, ng, tile, ncomp, posx, posy, ci, vector, halo, bonly, cap, *[a.ctypes.data_as(ip) for a in arrs])
    assert 0 <= cnt <= cap
    return [a[:cnt] for a in arrs]


@pytest.mark.parametrize("spec", [("scalar", cs.CENTER, None), ("scalar", cs.CORNER, None),
                                  ("vector", cs.NORTH, cs.EAST), ("vector", cs.EAST, cs.NORTH), ("edge", cs.NORTH, cs.EAST)])
def test_cxx_halo_tables_match_numpy_builder(built, spec):
    """The C++ table builder shipped in the CUDA library (csrc/halo.cu) against the NumPy one."""
    libpath = built.LIB
    if not os.path.exists(libpath):
        pytest.skip("CUDA library not built (no nvcc here)")
    lib = C.CDLL(libpath)
    kind, px, py = spec
    n, ng = 10, 3
    bonly = kind == "edge"
    tabs = cs.build_tables(n, ng, px, py, "vector" if py is not None else "scalar", None, bonly)
    NI = ((5 + (n + 2 * ng + 1)) + 7) // 8 * 8
    for tile in range(1, 7):
        for ci in range(2 if py is not None else 1):
            d, st, sc, s, sg = _lib_tables(lib, n, ng, tile, 2 if py is not None else 1, px, py if py is not None else 0, ci,
                                           1, ng, int(bonly))
            tb = tabs[tile][ci]
            pos = [px, py][ci]
            # convert the NumPy native-plane indices to the padded device plane
            def to_padded(flat, p):
                nj, ni = cs.plane_shape(n, ng, p)
                j, i = np.divmod(flat, ni)
                return j * NI + i + 5
            ref = {}
            for k in range(tb.dst.size):
                psrc = [px, py][tb.src_comp[k]] if py is not None else px
                ref[int(to_padded(tb.dst[k], pos))] = (int(tb.src_tile[k]), int(tb.src_comp[k]), int(to_padded(tb.src[k], psrc)), int(tb.sign[k]))
            got = {int(d[k]): (int(st[k]), int(sc[k]), int(s[k]), int(sg[k])) for k in range(d.size)}
            assert got == ref, (tile, ci)


def test_set_eta_l79_levels():
    """tools/fv_eta.F90 set_eta for km = 79 (var_hi, ptop = 1 Pa, stretch 1.03, pint = 100 hPa): structural properties of the
    restated generator (no golden table exists in the reference: the levels are computed at run time)."""
    from gfdl_atmos_cubed_sphere_b200 import init_state as I
    ak, bk, ks = I.set_eta_var_hi(79)
    assert ak.shape == (80,) and bk.shape == (80,)
    assert ak[0] == 1.0 and bk[0] == 0.0 and ak[-1] == 0.0 and bk[-1] == 1.0
    assert np.all(bk[:ks + 1] == 0.0) and np.all(np.diff(bk[ks:]) > 0.0)          # pure pressure above pint, sigma-like below
    for ps in (1.0e5, 7.0e4, 5.0e4):
        assert np.all(np.diff(ak + bk * ps) > 0.0)
    p = ak + bk * 1.0e5
    assert abs(p[ks] - 100.0e2) / 100.0e2 < 0.15                                  # pint is snapped to an interface near 100 hPa
    dz = 287.05 / 9.80665 * 270.0 * np.diff(np.log(p))                            # isothermal thickness used by var_hi
    assert 15.0 < dz[-1] < 60.0 and dz[-1] < dz[-2] < dz[-3]                       # thin, stretching layers at the surface


def test_set_eta_l127_levels():
    """tools/fv_eta.F90 set_eta for km = 127, default npz_type (var_gfs, ptop = 1 Pa, stretch 1.028, pint = 75 hPa): structural
    properties of the restated generator (the levels are computed at run time; the reference holds no table for this branch)."""
    from gfdl_atmos_cubed_sphere_b200 import init_state as I
    ak, bk, ks = I.set_eta_var_gfs(127)
    assert ak.shape == (128,) and bk.shape == (128,)
    assert ak[0] == 1.0 and bk[0] == 0.0 and ak[-1] == 0.0 and bk[-1] == 1.0
    assert np.all(bk[:ks + 1] == 0.0) and np.all(np.diff(bk[ks:]) > 0.0)
    for ps in (1.0e5, 7.0e4, 5.0e4):
        assert np.all(np.diff(ak + bk * ps) > 0.0)
    p = ak + bk * 1.0e5
    assert abs(p[ks] - 75.0e2) / 75.0e2 < 0.15
    dz = 287.05 / 9.80665 * 270.0 * np.diff(np.log(p))
    assert 10.0 < dz[-1] < 40.0 and dz[-1] < dz[-2] < dz[-3]
    assert np.allclose(dz[-26:-1] - dz[-25:], dz[-26] - dz[-25], rtol=1e-3)          # k_inc = 25 layers of linearly growing thickness
    a2, b2 = I.model_levels(127)
    assert np.array_equal(a2, ak) and np.array_equal(b2, bk)


def test_omega_unit_vectors():
    """ec1, ec2 (get_center_vect, fv_grid_utils.F90:1738-1779) and en1, en2 (:629-642), the unit vectors of the omega diagnostic
    adv_pe: unit length, tangent to the sphere at the cell centre / normal to the great circle through the edge's end points,
    ec1 . ec2 = cos_sg(5) (:355-356), zero in the corner ghost blocks (:1757-1760)."""
    from gfdl_atmos_cubed_sphere_b200 import grid as G
    n = 10
    tiles = G.make_cubed_sphere(n)[0]
    for t in (0, 3):
        a = tiles[t].arr
        e1, e2, n1, n2 = a["ec1"], a["ec2"], a["en1"], a["en2"]
        assert e1.shape == (n + 6, n + 6, 3) and n1.shape == (n + 1, n, 3) and n2.shape == (n, n + 1, 3)
        inner = (slice(3, -3), slice(3, -3))
        for e in (e1, e2):
            assert np.allclose(np.linalg.norm(e[inner], axis=-1), 1.0, atol=1e-14)
        assert np.abs(e1[0, 0]).max() == 0.0 and np.abs(e2[-1, -1]).max() == 0.0
        assert np.allclose((e1 * e2).sum(-1)[inner], a["cos_sg"][4][inner], atol=1e-14)
        lon, lat = a["agrid"][0][inner], a["agrid"][1][inner]
        pc = np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)
        assert np.abs((e1[inner] * pc).sum(-1)).max() < 1e-2 and np.abs((e2[inner] * pc).sum(-1)).max() < 1e-2   # tangent (to grid accuracy)
        glon, glat = a["grid"][0][3:-3, 3:-3], a["grid"][1][3:-3, 3:-3]                  # corners (1:npx, 1:npy)
        g3 = np.stack([np.cos(glat) * np.cos(glon), np.cos(glat) * np.sin(glon), np.sin(glat)], axis=-1)
        assert np.allclose(np.linalg.norm(n1, axis=-1), 1.0, atol=1e-14) and np.allclose(np.linalg.norm(n2, axis=-1), 1.0, atol=1e-14)
        assert np.abs((n1 * g3[:, :-1]).sum(-1)).max() < 1e-14 and np.abs((n1 * g3[:, 1:]).sum(-1)).max() < 1e-14   # normal to both end points
        assert np.abs((n2 * g3[:-1, :]).sum(-1)).max() < 1e-14 and np.abs((n2 * g3[1:, :]).sum(-1)).max() < 1e-14
