"""CPU, world_size 2, gloo: the N>1 host logic (face partition, per-peer message layout, canonical
message order, unpack-with-sign) gives exactly the single-process NumPy exchange."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out):
    sys.path.insert(0, ROOT)
    from gfdl_atmos_cubed_sphere_b200 import cubed_sphere as cs, parallel as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ng = 3
    rng = np.random.default_rng(7)
    full = {"c": [rng.standard_normal((2, n + 6, n + 6)) for _ in range(6)],
            "u": [rng.standard_normal((2, n + 7, n + 6)) for _ in range(6)],
            "v": [rng.standard_normal((2, n + 6, n + 7)) for _ in range(6)],
            "b": [rng.standard_normal((2, n + 7, n + 7)) for _ in range(6)]}
    ref = {k: [a.copy() for a in v] for k, v in full.items()}
    ex = cs.Exchanger(n, ng)
    ex.scalar(ref["c"], cs.CENTER); ex.scalar(ref["b"], cs.CORNER); ex.pair(ref["u"], ref["v"], cs.NORTH, cs.EAST, kind="vector")
    dx = P.DistributedExchanger(n, ng, rank, world)
    mine = {k: {t: full[k][t - 1] for t in dx.my} for k in full}
    dx.exchange({t: [mine["c"][t]] for t in dx.my}, cs.CENTER)
    dx.exchange({t: [mine["b"][t]] for t in dx.my}, cs.CORNER)
    dx.exchange({t: [mine["u"][t], mine["v"][t]] for t in dx.my}, cs.NORTH, cs.EAST, kind="vector")
    ok = all(np.array_equal(mine[k][t], ref[k][t - 1]) for k in full for t in dx.my)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


def test_partition():
    sys.path.insert(0, ROOT)
    from gfdl_atmos_cubed_sphere_b200 import parallel as P
    for world in (1, 2, 3, 4, 6, 8):
        owned = [t for r in range(world) for t in P.tiles_of_rank(r, world)]
        assert sorted(owned) == [1, 2, 3, 4, 5, 6]
        m = P.tile_rank_map(world)
        for r in range(world):
            assert all(m[t - 1] == r for t in P.tiles_of_rank(r, world))
    assert P.tiles_of_rank(7, 8) == [] and P.tiles_of_rank(6, 8) == []


@pytest.mark.parametrize("world", [2])
def test_gloo_halo_exchange_world2(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, 10, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
