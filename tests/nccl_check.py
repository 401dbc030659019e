"""Multi-rank check (run under torchrun on >= 2 GPUs; collected by tests/test_multi_gpu_nccl.py): faces spread over ranks with
the NCCL halo exchange must reproduce the single-process 6-face result BIT FOR BIT (same kernels, same tables), for the acoustic
loop and for tracer_2d's all-reduce(max).
  FV3_CHECK_FLAGSET=A|B     flag set
  FV3_CHECK_ATTACH=1        the communicator is created by the caller (ncclCommInitRank through ctypes) and handed to the
                            library with fv3_comm_attach instead of fv3_comm_init"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from gfdl_atmos_cubed_sphere_b200 import abi, Case, CudaCube, tiles_of_rank, tile_rank_map

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
case = Case(24, 6, os.environ.get("FV3_CHECK_FLAGSET", "A"), state="baroclinic")
lib = abi.load_library()
my = tiles_of_rank(rank, world)
cube = CudaCube.for_rank(case, rank, world, device=local)
if os.environ.get("FV3_CHECK_ATTACH") == "1":
    # a communicator the CALLER owns: created here with NCCL's own API, borrowed by the library
    class NcclId(C.Structure):
        _fields_ = [("internal", C.c_char * 128)]
    nccl = C.CDLL("libnccl.so.2")
    uid = NcclId()
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
        idbuf.copy_(torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    active = min(world, 6)
    if cube is not None:
        C.memmove(C.byref(uid), bytes(idbuf.cpu().numpy().tobytes()), 128)
        comm = C.c_void_p()
        nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, NcclId, C.c_int]
        assert nccl.ncclCommInitRank(C.byref(comm), active, uid, rank) == 0
        fn = lib[0].fv3_comm_attach
        fn.restype = C.c_int
        rc = fn(cube.ctxs, len(my), comm, C.c_int(rank), (C.c_int * 6)(*tile_rank_map(world)))
        assert rc == 0, cube.eng[my[0]].last_error()
else:
    CudaCube.attach_nccl(cube, rank, world)
ok = True
if cube is not None:
    cube.dyn_core(1200.0, 3)
    # reference: all 6 faces in one process on this rank's GPU
    ref = CudaCube(case, device=local)
    ref.dyn_core(1200.0, 3)
    for t in my:
        for f in ("U", "V", "W", "PT", "DELP", "DELZ", "MFX", "MFY", "CX", "CY"):
            a, b = cube.eng[t].get(f), ref.eng[t].get(f)
            if not np.array_equal(a, b):
                ok = False
                print(f"rank {rank} tile {t} field {f} differs: max {np.abs(a-b).max():.3e}")
    # tracer_2d: the CFL maximum is reduced over the ranks with ncclAllReduce(max) -- must equal the single-process reduction
    cm_d, cm_r = (C.c_double * 6)(), (C.c_double * 6)()
    for cb in (cube, ref):
        for t in cb.tiles:
            e = cb.eng[t]
            e.put("WORK_Q", 1.0 + 0.01 * e.get("PT")); e.put("DP1", case.states[t - 1]["delp"])
    fn = lib[0].fv3_tracer_2d
    fn.restype = C.c_int
    assert fn(cube.ctxs, len(my), C.c_int(8), cm_d) == 0, cube.eng[my[0]].last_error()
    assert fn(ref.ctxs, 6, C.c_int(8), cm_r) == 0
    if list(cm_d) != list(cm_r):
        ok = False
        print(f"rank {rank}: reduced cmax differs", list(cm_d), list(cm_r))
    for t in my:
        a, b = cube.eng[t].get("WORK_Q"), ref.eng[t].get("WORK_Q")
        if not np.array_equal(a, b):
            ok = False
            print(f"rank {rank} tile {t} tracer differs: max {np.abs(a-b).max():.3e}")
    # which exchange carried the data: peer-mapped (CUDA IPC arenas, no NCCL kernel on the data path) or NCCL send / recv
    perr = C.c_int(0)
    fn = lib[0].fv3_halo_p2p_status
    fn.restype = C.c_int
    on = fn(cube.eng[my[0]].ctx, C.byref(perr))
    want = os.environ.get("FV3_CHECK_EXPECT_P2P")
    print(f"rank {rank}: halo exchange = {'peer-mapped' if on == 1 else 'nccl'}, arrival timeout flag = {perr.value}")
    if perr.value != 0 or (want is not None and on != int(want)):
        ok = False
    if os.environ.get("FV3_CHECK_ATTACH") == "1":
        cube.close()      # must NOT destroy the borrowed communicator ...
        assert nccl.ncclCommDestroy(comm) == 0   # ... so the owner can
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("NCCL halo exchange + all-reduce(max) == single-process result:", bool(flag.item()))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
