"""Multi-rank check (run under torchrun on >=2 GPUs): faces spread over ranks with NCCL halo exchange
must reproduce the single-process 6-face result bit-for-bit (same kernels, same tables)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import harness as H
import bench as B
from gfdl_atmos_cubed_sphere_b200 import abi

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
case = H.Case(24, 6, os.environ.get("FV3_CHECK_FLAGSET", "A"), state="baroclinic")
lib = abi.load_library()
my = B.tiles_of_rank(rank, world)
cube = H.CudaCube(case, tiles=my, device=local, link=True) if len(my) > 1 else H.CudaCube(case, tiles=my, device=local, link=False)
if len(my) == 1:
    assert lib[0].fv3_cube_link(cube.ctxs, (C.c_int * 1)(*my), 1) == 0
idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    raw = C.create_string_buffer(128); assert lib[0].fv3_nccl_unique_id(raw) == 0
    idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
dist.broadcast(idbuf, 0)
tr = (C.c_int * 6)(*B.tile_rank_map(world))
rc = lib[0].fv3_comm_init(cube.ctxs, len(my), C.c_char_p(bytes(idbuf.cpu().numpy().tobytes())), min(world, 6), rank, tr)
assert rc == 0, cube.eng[my[0]].last_error()
cube.dyn_core(1200.0, 3)
# reference: all 6 faces in one process on this rank's GPU
ref = H.CudaCube(case, device=local)
ref.dyn_core(1200.0, 3)
ok = True
for t in my:
    for f in ("U", "V", "W", "PT", "DELP", "DELZ"):
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
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("NCCL halo exchange + all-reduce(max) == single-process result:", bool(flag.item()))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
