"""Device time of one fv_dynamics step (fv3_fv_dynamics: entry conversion, k_split x {dyn_core, tracer_2d, vertical remap}, omega
filter) over the full C{res}L{npz} cube on one GPU, next to the dyn_core call alone.
   usage: python profiles/prof_fv_dynamics.py [res] [npz] [nq] [k_split] [n_split]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from gfdl_atmos_cubed_sphere_b200 import Case, CudaCube

res = int(sys.argv[1]) if len(sys.argv) > 1 else 384
npz = int(sys.argv[2]) if len(sys.argv) > 2 else 79
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1
k_split = int(sys.argv[4]) if len(sys.argv) > 4 else 1
n_split = int(sys.argv[5]) if len(sys.argv) > 5 else 8
case = Case(res, npz, "A", state="baroclinic")
cube = CudaCube(case)
cube.set_num_tracers(nq)
kappa = case.consts["kappa"]
for t in cube.tiles:                     # temperature on entry, tracers = smooth positive fields
    e = cube.eng[t]
    pt, delp = e.get("PT"), e.get("DELP")
    p = case.ak[0] + np.cumsum(delp, axis=0) - 0.5 * delp
    T0 = pt * p ** kappa
    e.put("PT", T0)
    for iq in range(nq):
        e.call("select_tracer", iq); e.put("WORK_Q", 1.0e-3 * (iq + 1) * (p / 1.0e5) ** 2)
    e.call("select_tracer", 0)
bdt = 225.0 * k_split
lib = cube.lib[0]


def timed(fn, reps=3):
    out = []
    for r in range(reps + 1):
        lib.fv3_timer_start(cube.ctxs, len(cube.tiles))
        fn()
        ms = C.c_double(0)
        lib.fv3_timer_stop(cube.ctxs, len(cube.tiles), C.byref(ms))
        if r:
            out.append(ms.value)
    return sum(out) / len(out)


t_dyn = timed(lambda: cube.dyn_core(225.0, n_split))
t_all = timed(lambda: cube.fv_dynamics(bdt, k_split, n_split, 9, 9, -9, 9, 8, 1))
# per-stage device time of one more step (stage timers on: CUDA events around every stage of face 1)
e1 = cube.eng[cube.tiles[0]]
lib.fv3_stage_timers(e1.ctx, 1)
cube.fv_dynamics(bdt, k_split, n_split, 9, 9, -9, 9, 8, 1)
stages = {}
for nm in ("C_SW", "D_SW", "UPDATE_DZ_C", "UPDATE_DZ", "Riem_Solver_C", "Riem_Solver3", "PG_C", "PG_D", "REMAP", "OMEGA", "PT_TO_THETA", "DEL2_CUBED"):
    ms, calls = C.c_double(0), C.c_longlong(0)
    if lib.fv3_stage_time_ms(e1.ctx, nm.encode(), C.byref(ms), C.byref(calls)) == 0 and calls.value:
        stages[nm] = (ms.value, calls.value)
lib.fv3_stage_timers(e1.ctx, 0)
cells = 6 * res * res * npz
print(f"C{res}L{npz}, 6 faces on one GPU, nq = {nq}, k_split = {k_split}, n_split = {n_split}")
print(f"  dyn_core alone ({n_split} substeps)            {t_dyn:8.1f} ms")
print(f"  fv_dynamics (dyn_core + tracer_2d + remap + omega) {t_all:8.1f} ms = {t_all / k_split:.1f} ms per k_split iteration "
      f"= {cells * n_split * k_split / (t_all * 1e-3):.3e} cell-updates/s")
print("  stages of face 1 in one fv_dynamics step (ms total, calls):", {k: (round(v[0], 2), v[1]) for k, v in stages.items()})
