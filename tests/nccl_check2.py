import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import harness as H
import bench as B
from gfdl_atmos_cubed_sphere_b200 import abi, cubed_sphere as cs
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
case = H.Case(12, 2, "A", state="baroclinic")
lib = abi.load_library()
my = B.tiles_of_rank(rank, world)
cube = H.CudaCube(case, tiles=my, device=local, link=True)
idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    raw = C.create_string_buffer(128); assert lib[0].fv3_nccl_unique_id(raw) == 0
    idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
dist.broadcast(idbuf, 0)
tr = (C.c_int * 6)(*B.tile_rank_map(world))
rc = lib[0].fv3_comm_init(cube.ctxs, len(my), C.c_char_p(bytes(idbuf.cpu().numpy().tobytes())), min(world, 6), rank, tr)
print(rank, "comm_init rc", rc, flush=True)
# reference exchange in numpy on the full cube
ex = cs.Exchanger(12, 3)
for grp, fields in (("DELP_PT", ("DELP", "PT")), ("UVW", ("U", "V", "W"))):
    st = {f: [case.states[t][f.lower()].copy() for t in range(6)] for f in fields}
    if grp == "UVW":
        ex.pair(st["U"], st["V"], cs.NORTH, cs.EAST); ex.scalar(st["W"])
    else:
        for f in fields: ex.scalar(st[f])
    rc = lib[0].fv3_halo_exchange(cube.ctxs, len(my), abi.HALO_ID[grp])
    for t in my: cube.eng[t].sync()
    print(rank, grp, "rc", rc, flush=True)
    for t in my:
        for f in fields:
            a = cube.eng[t].get(f); b = st[f][t - 1]
            print(rank, "tile", t, f, "equal" if np.array_equal(a, b) else f"DIFF n={np.sum(a != b)} of {a.size} nan={np.isnan(a).sum()}", flush=True)
dist.destroy_process_group()
