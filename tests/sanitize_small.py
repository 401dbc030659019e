"""Small end-to-end pass over every kernel family for `compute-sanitizer --tool memcheck` (not a pytest):
   compute-sanitizer --tool memcheck --error-exitcode 3 python tests/sanitize_small.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import harness as H
from gfdl_atmos_cubed_sphere_b200 import abi

# one face, several transport tiles (interior + frame launches), flag-set B: every damping kernel
case = H.Case(56, 3, "B", state="baroclinic", flags_override=dict(use_cond=1))
e = case.engine(abi.load_library(), 1)
case.load_state(e, 1)
e.call("c_sw", 5.0); e.call("d_sw", 10.0); e.sync(); e.close()
# full cube, non-hydrostatic, 2 substeps, then the tracer
for flags in (dict(), dict(hydrostatic=1), dict(use_cond=1, moist_kappa=1)):
    case = H.Case(16, 5, "A", state="baroclinic", flags_override=flags)
    gc = H.CudaCube(case)
    dp1 = {t: gc.eng[t].get("DELP") for t in gc.tiles}
    gc.dyn_core(800.0, 2)
    if not flags:
        for t in gc.tiles:
            gc.eng[t].put("WORK_Q", 1.0 + 0.01 * gc.eng[t].get("PT")); gc.eng[t].put("DP1", dp1[t])
        fn = gc.lib[0].fv3_tracer_2d; fn.restype = C.c_int
        assert fn(gc.ctxs, 6, C.c_int(8), None) == 0
    assert np.isfinite(gc.eng[1].get("U")).all()
    gc.close()
print("sanitize_small: done")
