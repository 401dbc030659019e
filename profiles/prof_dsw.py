"""Profiling driver: one face of C384L79, c_sw then d_sw (x2), for ncu launch lists / full captures.
   usage: python profiles/prof_dsw.py [res] [npz] [flagset]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H
from gfdl_atmos_cubed_sphere_b200 import abi
res = int(sys.argv[1]) if len(sys.argv) > 1 else 384
npz = int(sys.argv[2]) if len(sys.argv) > 2 else 79
fs = sys.argv[3] if len(sys.argv) > 3 else "A"
case = H.Case(res, npz, fs, state="baroclinic")
e = case.engine(abi.load_library(), 1)
case.load_state(e, 1)
dt = 225.0 / 8 * 384 / res
e.call("c_sw", 0.5 * dt)
for _ in range(2):
    e.call("d_sw", dt)
e.sync()
import ctypes as C
e.lib.fv3_stage_timers(e.ctx, 1)
for _ in range(3):
    e.call("d_sw", dt)
e.sync()
ms, calls = C.c_double(0), C.c_longlong(0)
e.lib.fv3_stage_time_ms(e.ctx, b"D_SW", C.byref(ms), C.byref(calls))
print("D_SW ms/call (one face, alone):", ms.value / max(calls.value, 1))
