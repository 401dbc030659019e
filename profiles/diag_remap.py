"""where does the CUDA remap differ from the oracle?  (diagnostic)  python profiles/diag_remap.py kord"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
import test_remap_gpu as T
kord = int(sys.argv[1]) if len(sys.argv) > 1 else 11
case, oc, gc = T._pair()
n = T.N
for mode, iv in [(0, 1), (1, 1), (1, -1), (1, -2), (2, 0), (0, 0), (1, 2)]:
    eo, eg = oc.eng[1], gc.eng[1]
    q0 = eo.get("WORK_Q")
    for e in (eo, eg):
        e.call("remap_work_q", mode, iv, kord, 1.0 if mode == 0 else 0.0)
    a = H.sub(eg, "WORK_Q", eg.get("WORK_Q"), 1, n, 1, n); b = H.sub(eo, "WORK_Q", eo.get("WORK_Q"), 1, n, 1, n)
    d = np.abs(a - b)
    k, j, i = np.unravel_index(np.argmax(d), d.shape)
    print(f"kord {kord} mode {mode} iv {iv}: max err {d.max() / np.abs(b).max():.2e} at k={k + 1} j={j + 1} i={i + 1}; columns differing: {(d.max(axis=0) > 1e-12 * np.abs(b).max()).sum()} of {n * n}; levels: {np.nonzero(d.max(axis=(1, 2)) > 1e-12 * np.abs(b).max())[0] + 1}")
    for e in (eo, eg):
        e.put("WORK_Q", q0)
