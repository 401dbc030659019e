"""Ad-hoc GPU shake-out (not a pytest): prints per-field parity errors for the main stages."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import harness as H
def show(title, res):
    print(title, " ".join(f"{k}={v:.1e}{'!!' if not (v < 1e-10) else ''}" for k, v in res.items()))
for fs in ("A", "B"):
    case = H.Case(16, 6, fs, state="baroclinic")
    oc = H.OracleCube(case); gc = H.CudaCube(case)
    oc.dyn_core(800.0, 2); gc.dyn_core(800.0, 2)
    for t in oc.tiles:
        show(f"{fs} tile {t}", H.compare(oc.eng[t], gc.eng[t], H.regions_state(case.bounds)))
    oc.close(); gc.close()
