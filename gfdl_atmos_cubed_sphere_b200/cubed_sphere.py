"""Cubed-sphere mosaic topology and halo-exchange index tables (host logic).

Restates, by geometry, what FMS ``mpp_define_mosaic`` / ``mpp_update_domains`` do for the
6-tile mosaic the reference declares in ``tools/fv_mp_mod.F90:498-546`` (12 contacts, halo
width ``ng=3`` ``:61,564``, ``symmetry=.true.`` ``:432,563``).  FMS itself is not part of the
reference checkout (RELEASE.md:7, FMS 2024.03), so the index/sign rules are derived from the
contact table:

* every tile is the square ``[0,n]^2`` in local continuous coordinates, cell ``(i,j)``
  spanning ``[i-1,i]x[j-1,j]``;
* across a contact the neighbour's coordinates are an affine map ``p' = M p + c`` with ``M`` a
  signed permutation (rotation by a multiple of 90 degrees);
* a field value at a halo point is the neighbour's value at the mapped point; staggered
  fields map by geometric edge/corner, vector pairs (D-grid ``u,v``; C-grid ``uc,vc``)
  transform with ``M^T`` (component swap + one sign flip at rotated contacts), SCALAR_PAIR
  metrics with ``|M^T|``.

The same tables drive (a) the NumPy exchange used by the grid generator and by the CPU-oracle
driver, and (b) are cross-checked against the C++ table builder inside the CUDA library
(``csrc/halo.cpp``) by ``tests/test_halo_tables.py``.
"""
from __future__ import annotations

import numpy as np

# position types: offset of the point of entity (i,j) from (i, j) in continuous coordinates
CENTER, CORNER, NORTH, EAST = 0, 1, 2, 3
_POS_OFF = {CENTER: (-0.5, -0.5), CORNER: (-1.0, -1.0), NORTH: (-0.5, -1.0), EAST: (-1.0, -0.5)}
# extra extent (+1 in i, +1 in j) of the native array of each position type
_POS_EXT = {CENTER: (0, 0), CORNER: (1, 1), NORTH: (0, 1), EAST: (1, 0)}

W, E, S, N = 0, 1, 2, 3

# tools/fv_mp_mod.F90:499-546 -- (tileA, edgeA, tileB, edgeB, reversed)
CONTACTS = [
    (1, E, 2, W, False),  # :499-502
    (1, N, 3, W, True),   # :503-506
    (1, W, 5, N, True),   # :507-510
    (1, S, 6, N, False),  # :511-514
    (2, N, 3, S, False),  # :515-518
    (2, E, 4, S, True),   # :519-522
    (2, S, 6, E, True),   # :523-526
    (3, E, 4, W, False),  # :527-530
    (3, N, 5, W, True),   # :531-534
    (4, N, 5, S, False),  # :535-538
    (4, E, 6, S, True),   # :539-542
    (5, E, 6, W, False),  # :543-546
]


def neighbours():
    """nbr[tile][edge] = (other tile, other edge, reversed)."""
    nbr = {t: {} for t in range(1, 7)}
    for a, ea, b, eb, rev in CONTACTS:
        nbr[a][ea] = (b, eb, rev)
        nbr[b][eb] = (a, ea, rev)
    return nbr


def _edge_frame(edge, n):
    """origin, tangent (along increasing index), outward normal of an edge."""
    if edge == W:
        return np.array([0.0, 0.0]), np.array([0.0, 1.0]), np.array([-1.0, 0.0])
    if edge == E:
        return np.array([float(n), 0.0]), np.array([0.0, 1.0]), np.array([1.0, 0.0])
    if edge == S:
        return np.array([0.0, 0.0]), np.array([1.0, 0.0]), np.array([0.0, -1.0])
    return np.array([0.0, float(n)]), np.array([1.0, 0.0]), np.array([0.0, 1.0])


def affine(edge_a, edge_b, rev, n):
    """(M, c) with p_B = M p_A + c for points near edge_a of A / edge_b of B."""
    oa, ta, na = _edge_frame(edge_a, n)
    ob, tb, nb_ = _edge_frame(edge_b, n)
    # p_A = oa + s ta + d na ;  p_B = ob + s' tb - d nb,  s' = s or n - s
    sgn = -1.0 if rev else 1.0
    # s = ta.(p-oa), d = na.(p-oa)
    M = sgn * np.outer(tb, ta) - np.outer(nb_, na)
    c = ob + (n * tb if rev else 0.0) - M @ oa
    return M, c


class HaloTable:
    """Gather table for one tile, one destination array of a (possibly paired) field."""

    __slots__ = ("dst", "src_tile", "src_comp", "src", "sign")

    def __init__(self, dst, src_tile, src_comp, src, sign):
        self.dst = dst            # flat index into the destination (nj, ni) plane
        self.src_tile = src_tile  # 1..6
        self.src_comp = src_comp  # 0 = x-array of the pair, 1 = y-array
        self.src = src            # flat index into the source plane
        self.sign = sign          # +1 / -1


def plane_shape(n, ng, pos):
    ex, ey = _POS_EXT[pos]
    return (n + 2 * ng + ey, n + 2 * ng + ex)  # (nj, ni)


def _flat(i, j, n, ng, pos):
    nj, ni = plane_shape(n, ng, pos)
    return (j - (1 - ng)) * ni + (i - (1 - ng))


def build_tables(n, ng, pos_x, pos_y=None, kind="scalar", halo=None, boundary_only=False):
    """Tables for all 6 tiles.

    pos_x / pos_y : position type of the single array, or of the (x, y) pair
    kind          : 'scalar' | 'vector' (sign flips) | 'pair' (SCALAR_PAIR, no sign)
    halo          : halo width to fill (default ng)
    boundary_only : mpp_get_boundary semantics -- fill only the points ON the north (x-array)
                    / east (y-array) edge of the tile from the neighbour (dyn_core.F90:1151-1163)
    Returns tables[tile] = [HaloTable for x] (+ [HaloTable for y]).
    """
    halo = ng if halo is None else halo
    nbr = neighbours()
    comps = [pos_x] if pos_y is None else [pos_x, pos_y]
    # component direction carried by each array of a pair: x-array holds the i-component
    out = {}
    for t in range(1, 7):
        tabs = []
        for ci, pos in enumerate(comps):
            ox, oy = _POS_OFF[pos]
            ex, ey = _POS_EXT[pos]
            ii = np.arange(1 - ng, n + ng + ex + 1)
            jj = np.arange(1 - ng, n + ng + ey + 1)
            I, J = np.meshgrid(ii, jj)          # (nj, ni)
            X = I + ox
            Y = J + oy
            dst_l, st_l, sc_l, src_l, sg_l = [], [], [], [], []
            for edge in (W, E, S, N):
                if boundary_only:
                    if ci == 0 and pos == NORTH and edge == N:
                        m = (Y == n) & (X > 0) & (X < n)
                    elif ci == 1 and pos == EAST and edge == E:
                        m = (X == n) & (Y > 0) & (Y < n)
                    else:
                        continue
                else:
                    inx = (X >= 0) & (X <= n)
                    iny = (Y >= 0) & (Y <= n)
                    if edge == W:
                        m = (X < 0) & (X >= -halo) & iny
                    elif edge == E:
                        m = (X > n) & (X <= n + halo) & iny
                    elif edge == S:
                        m = (Y < 0) & (Y >= -halo) & inx
                    else:
                        m = (Y > n) & (Y <= n + halo) & inx
                if not m.any():
                    continue
                tb, eb, rev = nbr[t][edge]
                M, c = affine(edge, eb, rev, n)
                P = np.stack([X[m], Y[m]], axis=0)
                Q = M @ P + c[:, None]
                # component carried by this array in A: unit vector e (i-comp for ci==0)
                if len(comps) == 2:
                    e = np.zeros(2)
                    e[ci] = 1.0
                    eb_vec = M @ e  # the same direction expressed in B's axes
                    cj = int(np.argmax(np.abs(eb_vec)))  # which B component
                    sign = float(np.sign(eb_vec[cj]))
                    if kind == "pair":
                        sign = 1.0
                    pos_b = comps[cj]
                else:
                    cj, sign, pos_b = 0, 1.0, pos
                bx, by = _POS_OFF[pos_b]
                Ib = np.rint(Q[0] - bx).astype(np.int64)
                Jb = np.rint(Q[1] - by).astype(np.int64)
                assert np.allclose(Ib + bx, Q[0]) and np.allclose(Jb + by, Q[1]), "staggering mismatch"
                exb, eyb = _POS_EXT[pos_b]
                assert Ib.min() >= 1 and Ib.max() <= n + exb and Jb.min() >= 1 and Jb.max() <= n + eyb
                dst_l.append(_flat(I[m], J[m], n, ng, pos))
                st_l.append(np.full(Ib.shape, tb, dtype=np.int32))
                sc_l.append(np.full(Ib.shape, cj, dtype=np.int32))
                src_l.append(_flat(Ib, Jb, n, ng, pos_b))
                sg_l.append(np.full(Ib.shape, sign))
            tabs.append(HaloTable(np.concatenate(dst_l), np.concatenate(st_l), np.concatenate(sc_l),
                                  np.concatenate(src_l), np.concatenate(sg_l)))
        out[t] = tabs
    return out


class Exchanger:
    """NumPy halo exchange over 6 tiles (used by the grid generator and the oracle driver)."""

    def __init__(self, n, ng=3):
        self.n, self.ng = n, ng
        self._cache = {}

    def tables(self, pos_x, pos_y=None, kind="scalar", halo=None, boundary_only=False):
        key = (pos_x, pos_y, kind, halo, boundary_only)
        if key not in self._cache:
            self._cache[key] = build_tables(self.n, self.ng, pos_x, pos_y, kind, halo, boundary_only)
        return self._cache[key]

    def scalar(self, arrs, pos=CENTER, halo=None):
        """arrs: list of 6 arrays (..., nj, ni) updated in place."""
        tabs = self.tables(pos, halo=halo)
        flat = [a.reshape(a.shape[:-2] + (-1,)) for a in arrs]
        new = []
        for t in range(1, 7):
            tb = tabs[t][0]
            vals = np.empty(flat[t - 1].shape[:-1] + (tb.dst.size,))
            for s in range(1, 7):
                m = tb.src_tile == s
                if m.any():
                    vals[..., m] = flat[s - 1][..., tb.src[m]]
            new.append(vals)
        for t in range(1, 7):
            flat[t - 1][..., tabs[t][0].dst] = new[t - 1]

    def pair(self, xs, ys, pos_x, pos_y, kind="vector", boundary_only=False):
        """xs, ys: lists of 6 arrays; D-grid (u,v): NORTH,EAST; C-grid (uc,vc): EAST,NORTH."""
        tabs = self.tables(pos_x, pos_y, kind, boundary_only=boundary_only)
        fx = [a.reshape(a.shape[:-2] + (-1,)) for a in xs]
        fy = [a.reshape(a.shape[:-2] + (-1,)) for a in ys]
        src = (fx, fy)
        new = []
        for t in range(1, 7):
            pr = []
            for ci in range(2):
                tb = tabs[t][ci]
                vals = np.empty(src[ci][t - 1].shape[:-1] + (tb.dst.size,))
                for s in range(1, 7):
                    for cj in range(2):
                        m = (tb.src_tile == s) & (tb.src_comp == cj)
                        if m.any():
                            vals[..., m] = src[cj][s - 1][..., tb.src[m]] * tb.sign[m]
                pr.append(vals)
            new.append(pr)
        for t in range(1, 7):
            for ci in range(2):
                src[ci][t - 1][..., tabs[t][ci].dst] = new[t - 1][ci]
