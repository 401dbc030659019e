"""Face -> rank partition and the message protocol of the multi-rank halo exchange (host logic).

The CUDA library implements this protocol with NCCL send/recv (csrc/halo.cu); this module is
the same protocol on torch.distributed tensors, so it can be exercised with the ``gloo`` backend
on CPU (tests/test_distributed_gloo.py) -- partition, per-peer message layout, canonical message
order and unpack-with-sign are identical.

Reference: one rank per (sub-)tile with FMS group updates, tools/fv_mp_mod.F90:276-641, :646-874.
"""
from __future__ import annotations

import numpy as np

from . import cubed_sphere as cs


def tiles_of_rank(rank: int, world: int):
    """Faces 1..6 dealt round-robin over min(world, 6) active ranks (ranks >= 6 idle)."""
    active = min(world, 6)
    if rank >= active:
        return []
    return [t for t in range(1, 7) if (t - 1) % active == rank]


def tile_rank_map(world: int):
    active = min(world, 6)
    return [(t - 1) % active for t in range(1, 7)]


class DistributedExchanger:
    """Halo exchange of the faces owned by this rank; off-rank faces through dist.send/recv."""

    def __init__(self, n, ng, rank, world):
        self.n, self.ng, self.rank, self.world = n, ng, rank, world
        self.my = tiles_of_rank(rank, world)
        self.owner = tile_rank_map(world)
        self._cache = {}

    def _tables(self, pos_x, pos_y, kind, boundary_only=False):
        key = (pos_x, pos_y, kind, boundary_only)
        if key not in self._cache:
            self._cache[key] = cs.build_tables(self.n, self.ng, pos_x, pos_y, kind, None, boundary_only)
        return self._cache[key]

    def exchange(self, fields, pos_x, pos_y=None, kind="scalar", boundary_only=False):
        """fields: {tile: [x_array] or [x_array, y_array]} for the tiles of this rank (updated in place)."""
        import torch
        import torch.distributed as dist
        tabs = self._tables(pos_x, pos_y, kind, boundary_only)
        ncomp = 1 if pos_y is None else 2
        flat = {t: [a.reshape(a.shape[:-2] + (-1,)) for a in fields[t]] for t in self.my}
        # messages: (my tile, remote tile) pairs, canonical order per peer = (tile of lower rank, tile of higher rank)
        msgs = []
        for t in self.my:
            for r in range(1, 7):
                rr = self.owner[r - 1]
                if rr == self.rank or r in self.my:
                    continue
                need = any(((tabs[t][ci].src_tile == r).any()) for ci in range(ncomp))
                if need:
                    lo, hi = (t, r) if self.rank < rr else (r, t)
                    msgs.append((rr, lo, hi, t, r))
        msgs.sort()
        reqs, recvs = [], []
        for rr, _, _, t, r in msgs:
            # pack what tile r needs from my tile t: r's table entries with src_tile == t, in (ci, sc) order
            parts = []
            for ci in range(ncomp):
                tb = tabs[r][ci]
                for sc in range(ncomp):
                    m = (tb.src_tile == t) & (tb.src_comp == sc)
                    if m.any():
                        parts.append(flat[t][sc][..., tb.src[m]].reshape(-1))
            sendbuf = torch.from_numpy(np.ascontiguousarray(np.concatenate(parts)))
            nrecv = 0
            lead = int(np.prod(flat[t][0].shape[:-1]))
            for ci in range(ncomp):
                tb = tabs[t][ci]
                nrecv += int((tb.src_tile == r).sum()) * lead
            recvbuf = torch.empty(nrecv, dtype=torch.float64)
            reqs.append(dist.isend(sendbuf, rr))
            reqs.append(dist.irecv(recvbuf, rr))
            recvs.append((t, r, recvbuf, sendbuf))
        # local gathers (sources are compute-domain points, destinations halo points: disjoint)
        new = []
        for t in self.my:
            for ci in range(ncomp):
                tb = tabs[t][ci]
                for s in self.my:
                    for sc in range(ncomp):
                        m = (tb.src_tile == s) & (tb.src_comp == sc)
                        if m.any():
                            new.append((t, ci, tb.dst[m], flat[s][sc][..., tb.src[m]] * tb.sign[m]))
        for rq in reqs:
            rq.wait()
        for t, ci, d, v in new:
            flat[t][ci][..., d] = v
        for t, r, recvbuf, _ in recvs:
            buf = recvbuf.numpy()
            off = 0
            lead_shape = flat[t][0].shape[:-1]
            lead = int(np.prod(lead_shape))
            for ci in range(ncomp):
                tb = tabs[t][ci]
                for sc in range(ncomp):
                    m = (tb.src_tile == r) & (tb.src_comp == sc)
                    k = int(m.sum())
                    if k:
                        vals = buf[off:off + k * lead].reshape(lead_shape + (k,))
                        flat[t][ci][..., tb.dst[m]] = vals * tb.sign[m]
                        off += k * lead
