"""where do the fp32-sweep results differ from the fp64 ones?  (diagnostic; python profiles/diag_fp32.py [n] [npz] [substeps])"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
npz = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 1
case = H.Case(n, npz, "A", state="baroclinic")
out = []
for mode in (0, 1):
    gc = H.CudaCube(case)
    gc.set_transport_fp32(mode)
    gc.dyn_core(225.0 * ns, ns)
    out.append({t: {f: gc.eng[t].get(f) for f in ("DELP", "PT", "W", "U", "V", "DELZ")} for t in gc.tiles})
    dims = {f: gc.eng[1].dims(f) for f in ("DELP", "PT", "W", "U", "V", "DELZ")}
    gc.close()
for f in ("DELP", "PT", "W", "U", "V", "DELZ"):
    ilo, ni, jlo, nj, nk, kmid = dims[f]
    for t in (1, 2, 3):
        a, b = out[1][t][f], out[0][t][f]
        sl = (slice(None), slice(1 - jlo, 1 - jlo + n), slice(1 - ilo, 1 - ilo + n))
        d = np.abs(a[sl] - b[sl])
        k, j, i = np.unravel_index(np.argmax(d), d.shape)
        print(f"{f:5s} tile {t}: max |d| {d.max():.3e} (rel to max {np.abs(b[sl]).max():.3e}: {d.max() / np.abs(b[sl]).max():.2e}) at k={k} j={j + 1} i={i + 1}; "
              f"per level max: {' '.join(f'{x:.1e}' for x in d.max(axis=(1, 2)))}")
