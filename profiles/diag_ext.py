"""diagnostic: hydrostatic d_ext > 0, CUDA vs oracle after n substeps in ONE dyn_core call"""
import sys; sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, harness as H
for beta in (0.0, 0.4):
    for ns in (1, 2, 4):
        over = dict(hydrostatic=1, d_ext=0.02, beta=beta)
        case = H.Case(24, 8, "A", state="baroclinic", flags_override=over)
        oc, gc = H.OracleCube(case), H.CudaCube(case)
        oc.dyn_core(450.0 * ns, ns); gc.dyn_core(450.0 * ns, ns)
        reg = {"U": (1, 24, 1, 25), "V": (1, 25, 1, 24), "DELP": (1, 24, 1, 24), "VT": (1, 25, 1, 25)}
        print("beta", beta, "n_split", ns, {k: f"{v:.1e}" for k, v in H.compare(oc.eng[1], gc.eng[1], reg).items()})
        oc.close(); gc.close()
