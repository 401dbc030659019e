"""Per-stage device time of one acoustic substep on ONE face of C{res}L{npz}, each stage timed alone
(CUDA events on the launch stream, nothing else in flight).  The stages run on the initial state in the
dyn_core order; a single face has no valid halo after the first substep, so this driver is for timing and
ncu launch lists only (parity lives in tests/).
   usage: python profiles/prof_stages.py [res] [npz] [flagset] [reps]      (FV3_TRANSPORT_FP32=1: fv3_set_transport_fp32)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H
from gfdl_atmos_cubed_sphere_b200 import abi

res = int(sys.argv[1]) if len(sys.argv) > 1 else 384
npz = int(sys.argv[2]) if len(sys.argv) > 2 else 79
fs = sys.argv[3] if len(sys.argv) > 3 else "A"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
case = H.Case(res, npz, fs, state="baroclinic")
e = case.engine(abi.load_library(), 1)
if os.environ.get("FV3_TRANSPORT_FP32") == "1":
    e.lib.fv3_set_transport_fp32(e.ctx, 1)
dt = 225.0 / 8 * 384 / res
F = abi.FIELD_ID
one = (C.c_void_p * 1)(e.ctx)
seq = [("gz_init", ()), ("C_SW", ("c_sw", 0.5 * dt)), ("copy", ("copy_field", F["ZH"], F["GZ"])),
       ("UPDATE_DZ_C", ("update_dz_c", 0.5 * dt)), ("Riem_Solver_C", ("riem_solver_c", 0.5 * dt)), ("PG_C", ("p_grad_c", 0.5 * dt)),
       ("D_SW", ("d_sw", dt)), ("UPDATE_DZ", ("update_dz_d", dt)), ("Riem_Solver3", ("riem_solver3", dt, 0)),
       ("pk3_halo", ("pk3_halo",)), ("gz_from_zh", ("gz_from_zh",)), ("PG_D", ("nh_p_grad", dt))]
tot = {}
for rep in range(reps + 1):
    case.load_state(e, 1)
    e.sync()
    for name, call in seq:
        if not call:
            e.call("gz_init"); e.sync(); continue
        e.lib.fv3_timer_start(one, 1)
        e.call(*call)
        ms = C.c_double(0)
        e.lib.fv3_timer_stop(one, 1, C.byref(ms))
        if rep > 0:
            tot[name] = tot.get(name, 0.0) + ms.value / reps
s = sum(tot.values())
print(f"one face C{res}L{npz} flag-set {fs}{' fp32 sweeps' if os.environ.get('FV3_TRANSPORT_FP32') == '1' else ''}: {s:.3f} ms per substep (sum of stages, each alone)")
for k, v in tot.items():
    print(f"  {k:14s} {v:8.3f} ms {100 * v / s:5.1f}%")
