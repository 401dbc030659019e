"""The cubed sphere on the CUDA library: configuration (`Case`) and the multi-face driver (`CudaCube`).

`Case` bundles what `fv3_create` needs for every face -- the gnomonic grid and its metric terms (grid.py), the hybrid levels
and an initial state (init_state.py), the flag set (`fv_flags_type` subset, abi.FLAGSET_A / _B).  `CudaCube` creates one
`fv3_ctx` per face owned by this process, links them for the halo exchange (`fv3_cube_link`), attaches the library's NCCL
communicator when the faces are spread over ranks (`attach_nccl`, bootstrap through torch.distributed) and steps the
acoustic loop (`fv3_dyn_core`).  Reference roles: `fv_control_init` + `domain_decomp` (tools/fv_control.F90,
tools/fv_mp_mod.F90:276-641) and the `dyn_core` call of `fv_dynamics` (model/fv_dynamics.F90:478-486).
"""
from __future__ import annotations

import ctypes as C
import functools

from . import abi, grid as G, init_state as I
from .parallel import tiles_of_rank, tile_rank_map

FLAGSETS = {"A": abi.FLAGSET_A, "B": abi.FLAGSET_B}


@functools.lru_cache(maxsize=4)
def cube_grid(n):
    return G.make_cubed_sphere(n)


class Case:
    def __init__(self, n, npz, flagset="A", state="smooth", flags_override=None, state_kw=None):
        self.n, self.npz = n, npz
        self.tiles, self.bounds = cube_grid(n)
        self.ak, self.bk = I.model_levels(npz)   # npz = 79: the reference set_eta levels (var_hi)
        self.flags = dict(FLAGSETS[flagset]) if isinstance(flagset, str) else dict(flagset)
        if flags_override:
            self.flags.update(flags_override)
        if state == "smooth":
            self.states = I.smooth_state(self.tiles, self.bounds, npz)
        elif state == "cosine_bell":   # SW_DYNAMICS test case 1 (BASELINE config 1a): npz = 1, flags.sw_test_case = 1
            assert npz == 1
            self.flags["sw_test_case"] = 1
            self.states = I.cosine_bell(self.tiles, self.bounds, **(state_kw or {}))
        else:
            self.states = I.baroclinic_wave(self.tiles, self.bounds, npz, self.ak, self.bk)
        self.consts = G.CONSTANTS

    def engine(self, lib_prefix, tile=1, device=0):
        lib, prefix = lib_prefix
        return abi.Engine(lib, prefix, self.bounds, self.tiles[tile - 1], self.flags, self.npz, self.ak, self.bk,
                          self.ak[0], self.consts, tile=tile, device=device)

    def load_state(self, eng, tile=1, fields=("u", "v", "w", "pt", "delp", "q_con", "phis", "delz", "uc", "vc")):
        st = self.states[tile - 1]
        name = {"q_con": "QCON"}
        for f in fields:
            if f in st:
                eng.put(name.get(f, f.upper()), st[f])



class CudaCube:
    """Faces of the cube on the CUDA library (all faces of this process on one GPU)."""

    def __init__(self, case, tiles=(1, 2, 3, 4, 5, 6), device=0, link=True):
        self.case = case
        self.tiles = list(tiles)
        self.lib = abi.load_library()
        self.eng = {t: case.engine(self.lib, t, device) for t in self.tiles}
        for t in self.tiles:
            case.load_state(self.eng[t], t)
        self.ctxs = (C.c_void_p * len(self.tiles))(*[self.eng[t].ctx for t in self.tiles])
        if link and len(self.tiles) > 1:
            tl = (C.c_int * len(self.tiles))(*self.tiles)
            rc = self.lib[0].fv3_cube_link(self.ctxs, tl, len(self.tiles))
            if rc:
                raise RuntimeError(f"fv3_cube_link rc={rc}: {self.eng[self.tiles[0]].last_error()}")

    def dyn_core(self, bdt, n_split, graph=False, end_step=False):
        """fv3_dyn_core on the faces of this process; graph=True: FV3_DYN_GRAPH (captured once, replayed with one launch);
        end_step=True: FV3_DYN_END_STEP (last call of the k_split loop: omega diagnostic on the last substep)."""
        fn = self.lib[0].fv3_dyn_core
        fn.restype = C.c_int
        rc = fn(self.ctxs, len(self.tiles), C.c_double(bdt), C.c_int(n_split), C.c_int((1 if graph else 0) | (2 if end_step else 0)))
        if rc:
            raise RuntimeError(f"fv3_dyn_core rc={rc}: " + "; ".join(self.eng[t].last_error() for t in self.tiles))
        for t in self.tiles:
            self.eng[t].sync()

    def fv_dynamics(self, bdt, k_split, n_split, kord_mt=9, kord_wz=9, kord_tm=-9, kord_tr=9, hord_tr=0, nf_omega=1, graph=False,
                    sphum=-1, zvir=0.0):
        """fv3_fv_dynamics[_qv]: entry conversion, k_split x (dyn_core, tracer_2d, vertical remap), omega filter; pt = T in and out;
        sphum >= 0: that tracer is the specific humidity (zvir = rvgas / rdgas - 1)."""
        fn = self.lib[0].fv3_fv_dynamics_qv
        fn.restype = C.c_int
        rc = fn(self.ctxs, len(self.tiles), C.c_double(bdt), C.c_int(k_split), C.c_int(n_split), C.c_int(kord_mt), C.c_int(kord_wz),
                C.c_int(kord_tm), C.c_int(kord_tr), C.c_int(hord_tr), C.c_int(nf_omega), C.c_int(1 if graph else 0), C.c_int(sphum),
                C.c_double(zvir))
        if rc:
            raise RuntimeError(f"fv3_fv_dynamics rc={rc}: " + "; ".join(self.eng[t].last_error() for t in self.tiles))
        for t in self.tiles:
            self.eng[t].sync()

    def set_num_tracers(self, nq):
        """fv3_set_num_tracers on every face: nq tracer arrays; FV3_WORK_Q names the one chosen with select_tracer."""
        for t in self.tiles:
            self.eng[t].call("set_num_tracers", int(nq))

    def set_tracer_fill(self, on):
        """flagstruct%fill: fillz after each remapped tracer (fv3_set_tracer_fill on every face)"""
        for t in self.tiles:
            self.eng[t].call("set_tracer_fill", int(bool(on)))

    def select_tracer(self, iq):
        for t in self.tiles:
            self.eng[t].call("select_tracer", int(iq))

    def set_transport_fp32(self, on=True):
        """fv3_set_transport_fp32 on every face: fp32 PPM sweeps on the interior tiles of d_sw (BASELINE config 5)."""
        fn = self.lib[0].fv3_set_transport_fp32
        fn.restype = C.c_int
        for t in self.tiles:
            rc = fn(self.eng[t].ctx, C.c_int(1 if on else 0))
            if rc:
                raise RuntimeError(f"fv3_set_transport_fp32 rc={rc}")

    def del2_cubed(self, field, cd, nmax):
        fn = self.lib[0].fv3_del2_cubed_cube
        fn.restype = C.c_int
        rc = fn(self.ctxs, len(self.tiles), C.c_int(abi.FIELD_ID[field]), C.c_double(cd), C.c_int(nmax))
        if rc:
            raise RuntimeError(f"fv3_del2_cubed_cube rc={rc}: " + "; ".join(self.eng[t].last_error() for t in self.tiles))
        for t in self.tiles:
            self.eng[t].sync()

    def close(self):
        for e in self.eng.values():
            e.close()

    # ---- multi-rank: one process per GPU, faces dealt by parallel.tiles_of_rank ----------------------------------------
    @classmethod
    def for_rank(cls, case, rank, world, device=0):
        """The faces of `rank` (None for an idle rank: more than 6 ranks), linked for the exchange."""
        my = tiles_of_rank(rank, world)
        if not my:
            return None
        cube = cls(case, tiles=my, device=device, link=len(my) > 1)
        if len(my) == 1 and world > 1:   # a single face still needs its halo plan for the off-rank contacts
            rc = cube.lib[0].fv3_cube_link(cube.ctxs, (C.c_int * 1)(*my), 1)
            if rc:
                raise RuntimeError(f"fv3_cube_link rc={rc}: {cube.eng[my[0]].last_error()}")
        return cube

    @staticmethod
    def attach_nccl(cube, rank, world):
        """Create the library's own NCCL communicator over the active ranks (collective over torch.distributed's default
        group, which only carries the unique id).  `cube` is None on idle ranks -- they take part in the broadcast only."""
        import torch
        import torch.distributed as dist
        lib = abi.load_library()[0]
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            if lib.fv3_nccl_unique_id(raw) != 0:
                raise RuntimeError("fv3_nccl_unique_id failed (NCCL not loadable)")
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        if cube is None:
            return
        raw = bytes(idbuf.cpu().numpy().tobytes())
        tr = (C.c_int * 6)(*tile_rank_map(world))
        rc = lib.fv3_comm_init(cube.ctxs, len(cube.tiles), C.c_char_p(raw), min(world, 6), rank, tr)
        if rc:
            raise RuntimeError(f"fv3_comm_init rc={rc}: {cube.eng[cube.tiles[0]].last_error()}")
