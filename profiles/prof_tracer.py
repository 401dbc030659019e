"""Device time of fv3_tracer_2d (one tracer, all six faces of C{res}L{npz} on one GPU) after one dyn_core call.
   usage: python profiles/prof_tracer.py [res] [npz] [hord]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
res = int(sys.argv[1]) if len(sys.argv) > 1 else 384
npz = int(sys.argv[2]) if len(sys.argv) > 2 else 79
hord = int(sys.argv[3]) if len(sys.argv) > 3 else 8
case = H.Case(res, npz, "A", state="baroclinic")
gc = H.CudaCube(case)
lib = gc.lib[0]
dp1 = {t: gc.eng[t].get("DELP") for t in gc.tiles}
gc.dyn_core(225.0 * 384 / res, 8)
saved = {t: {f: gc.eng[t].get(f) for f in ("MFX", "MFY", "CX", "CY")} for t in gc.tiles}
fn = lib.fv3_tracer_2d; fn.restype = C.c_int
cm = (C.c_double * npz)()
best = None
for rep in range(3):
    for t in gc.tiles:
        e = gc.eng[t]
        e.put("WORK_Q", 1.0 + 0.01 * e.get("PT")); e.put("DP1", dp1[t])
        for f, a in saved[t].items():
            e.put(f, a)
        e.sync()
    t0 = time.perf_counter()
    assert fn(gc.ctxs, 6, C.c_int(hord), cm) == 0
    dt = time.perf_counter() - t0     # the call synchronises before it returns
    best = dt if best is None else min(best, dt)
ns = (1.0 + np.array(cm[:])).astype(int)
print(f"tracer_2d C{res}L{npz} hord {hord}, 6 faces: {best*1e3:.2f} ms per tracer; nsplt max {ns.max()}, sub-cycles summed over levels {ns.sum()} "
      f"=> {6*res*res*ns.sum()/best:.3e} cell-updates/s")
