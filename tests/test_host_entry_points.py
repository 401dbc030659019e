"""-m gpu: the entry points a Fortran shim would bind: the per-call host-buffer drop-ins fv3_c_sw_host,
fv3_d_sw_host, fv3_fv_tp_2d_host, fv3_riem_solver_c_host and the whole-state fv3_upload_state -> fv3_dyn_core ->
fv3_download_state.  Every buffer is a NumPy array in the reference's native Fortran extents (column-major, i fastest:
dyn_core.F90:436-447, 531-536, 762-772; fv_arrays.F90:1521-1563) -- none of them goes through Engine.put / Engine.get on
the CUDA side, so this is the ABI as the reference's call sites would use it.  Checker: the CPU oracle on the same inputs.
"""
import ctypes as C

import numpy as np
import pytest

import harness as H
from gfdl_atmos_cubed_sphere_b200 import abi

pytestmark = pytest.mark.gpu
TOL = 1e-12          # per call, horizontal operators (tolerances: header of tests/test_gpu_parity.py)
TOL_SOLVER = 1e-11   # outputs of a column solver
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _err(a, b):
    den = max(float(np.abs(b).max()), 1e-300)
    return float(np.abs(a - b).max()) / den


def _section(eng, name, arr, box):
    return H.sub(eng, name, arr, *box)


def _setup(n=24, npz=5, flagset="A", state="smooth", **over):
    case = H.Case(n, npz, flagset, state=state, flags_override=over or None)
    eo = case.engine(H.load_oracle(), 1)
    eg = case.engine(abi.load_library(), 1)
    case.load_state(eo, 1)
    return case, eo, eg


def test_c_sw_host_and_d_sw_host():
    case, eo, eg = _setup()
    lib = eg.lib
    st = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in case.states[0].items()}
    dt = 20.0
    # ---- c_sw: inputs delp, pt, u, v, w; everything else is output
    out = {f: np.zeros(eg.shape(f)) for f in ("DELPC", "PTC", "UC", "VC", "UA", "VA", "OMGA", "UT", "VT", "DIVGD")}
    fn = lib.fv3_c_sw_host
    fn.restype = C.c_int
    rc = fn(eg.ctx, _p(out["DELPC"]), _p(st["delp"]), _p(out["PTC"]), _p(st["pt"]), _p(st["u"]), _p(st["v"]), _p(st["w"]), _p(out["UC"]),
            _p(out["VC"]), _p(out["UA"]), _p(out["VA"]), _p(out["OMGA"]), _p(out["UT"]), _p(out["VT"]), _p(out["DIVGD"]), C.c_double(0.5 * dt))
    assert rc == 0, eg.last_error()
    eo.call("c_sw", 0.5 * dt)
    for f, box in H.regions_c_sw(case.bounds).items():
        assert _err(_section(eg, f, out[f], box), _section(eo, f, eo.get(f), box)) <= TOL, f
    # ---- d_sw on the oracle's c_sw results (host buffers again; in/out arrays are updated in place)
    io = {f: eo.get(f).copy() for f in ("DELP", "PT", "U", "V", "W", "UC", "VC", "UA", "VA", "DIVGD", "MFX", "MFY", "CX", "CY")}
    o2 = {f: np.zeros(eg.shape(f)) for f in ("CRX", "CRY", "XFX", "YFX")}
    fn = lib.fv3_d_sw_host
    fn.restype = C.c_int
    rc = fn(eg.ctx, _p(io["DELP"]), _p(io["PT"]), _p(io["U"]), _p(io["V"]), _p(io["W"]), _p(io["UC"]), _p(io["VC"]), _p(io["UA"]), _p(io["VA"]),
            _p(io["DIVGD"]), _p(io["MFX"]), _p(io["MFY"]), _p(io["CX"]), _p(io["CY"]), _p(o2["CRX"]), _p(o2["CRY"]), _p(o2["XFX"]), _p(o2["YFX"]),
            None, None, None, C.c_double(dt))
    assert rc == 0, eg.last_error()
    eo.call("d_sw", dt)
    got = dict(io); got.update(o2)
    for f, box in H.regions_d_sw(case.bounds).items():
        assert _err(_section(eg, f, got[f], box), _section(eo, f, eo.get(f), box)) <= TOL, f
    eo.close(); eg.close()


@pytest.mark.parametrize("hord,use_mfx", [(8, 0), (-5, 1)])
def test_fv_tp_2d_host(hord, use_mfx):
    case, eo, eg = _setup(n=30, npz=4)
    rng = np.random.default_rng(3)
    b = case.bounds
    g = case.tiles[0].arr
    shp = {f: eg.shape(f) for f in ("WORK_Q", "CRX", "CRY", "XFX", "YFX", "WORK_RAX", "WORK_RAY", "WORK_FX", "WORK_FY", "MFX", "MFY")}
    q = 1.0 + 0.3 * rng.random(shp["WORK_Q"])
    crx = 0.6 * (rng.random(shp["CRX"]) - 0.5); cry = 0.6 * (rng.random(shp["CRY"]) - 0.5)
    is_, ie, js, je, isd, ied, jsd, jed = b["is_"], b["ie"], b["js"], b["je"], b["isd"], b["ied"], b["jsd"], b["jed"]
    area = g["area"]
    xfx = crx * area[None, :, is_ - isd - 0:ie - isd + 2] * 0.9
    yfx = cry * area[None, js - jsd:je - jsd + 2, :] * 0.9
    ra_x = area[None, :, is_ - isd:ie - isd + 1] + xfx[:, :, :-1] - xfx[:, :, 1:]
    ra_y = area[None, js - jsd:je - jsd + 1, :] + yfx[:, :-1, :] - yfx[:, 1:, :]
    mfx = 50.0 * rng.random(shp["MFX"]); mfy = 50.0 * rng.random(shp["MFY"])
    arrs = {"WORK_Q": q, "CRX": crx, "CRY": cry, "XFX": xfx, "YFX": yfx, "WORK_RAX": ra_x, "WORK_RAY": ra_y, "MFX": mfx, "MFY": mfy}
    arrs = {k: np.ascontiguousarray(v) for k, v in arrs.items()}
    for f, a in arrs.items():
        assert a.shape == shp[f], (f, a.shape, shp[f])
        eo.put(f, a)
    eo.call("fv_tp_2d", 4, hord, use_mfx, 0, -1, 0.0)
    fx = np.zeros(shp["WORK_FX"]); fy = np.zeros(shp["WORK_FY"])
    fn = eg.lib.fv3_fv_tp_2d_host
    fn.restype = C.c_int
    rc = fn(eg.ctx, C.c_int(4), _p(arrs["WORK_Q"]), _p(arrs["CRX"]), _p(arrs["CRY"]), _p(arrs["XFX"]), _p(arrs["YFX"]), _p(arrs["WORK_RAX"]),
            _p(arrs["WORK_RAY"]), C.c_int(hord), _p(fx), _p(fy), _p(arrs["MFX"]) if use_mfx else None, _p(arrs["MFY"]) if use_mfx else None,
            None, C.c_int(-1), C.c_double(0.0))
    assert rc == 0, eg.last_error()
    assert _err(fx, eo.get("WORK_FX")) <= TOL and _err(fy, eo.get("WORK_FY")) <= TOL
    eo.close(); eg.close()


def test_riem_solver_c_host():
    case, eo, eg = _setup(n=16, npz=8, state="baroclinic")
    dt2 = 12.0
    # the C-grid state Riem_Solver_c works on: c_sw + gz + update_dz_c on the oracle, then both solvers from those buffers
    eo.call("c_sw", dt2); eo.call("gz_init")
    eo.call("update_dz_c", dt2)
    names = ("CAPPA", "PHIS", "OMGA", "PTC", "QCON", "DELPC", "GZ", "WS3")
    buf = {f: eo.get(f).copy() for f in names}
    pef = np.zeros(eg.shape("PKC"))
    fn = eg.lib.fv3_riem_solver_c_host
    fn.restype = C.c_int
    rc = fn(eg.ctx, C.c_double(dt2), _p(buf["CAPPA"]), _p(buf["PHIS"]), _p(buf["OMGA"]), _p(buf["PTC"]), _p(buf["QCON"]), _p(buf["DELPC"]),
            _p(buf["GZ"]), _p(pef), _p(buf["WS3"]))
    assert rc == 0, eg.last_error()
    eo.call("riem_solver_c", dt2)
    b = case.bounds
    box = (b["is_"] - 1, b["ie"] + 1, b["js"] - 1, b["je"] + 1)
    assert _err(_section(eg, "GZ", buf["GZ"], box), _section(eo, "GZ", eo.get("GZ"), box)) <= TOL_SOLVER
    assert _err(_section(eg, "PKC", pef, box), _section(eo, "PKC", eo.get("PKC"), box)) <= TOL_SOLVER
    eo.close(); eg.close()


def test_upload_state_dyn_core_download_state():
    """fv3_state_t of host pointers in (dyn_core.F90:94-98 dummy arguments) -> two acoustic substeps of the linked cube ->
    fv3_state_t of host pointers out, against the oracle driven through its own field vocabulary."""
    n, npz = 16, 6
    case = H.Case(n, npz, "A", state="baroclinic")
    oc = H.OracleCube(case)
    lib = abi.load_library()
    eng = {t: case.engine(lib, t) for t in range(1, 7)}     # contexts WITHOUT load_state: the state arrives through fv3_upload_state
    ctxs = (C.c_void_p * 6)(*[eng[t].ctx for t in range(1, 7)])
    assert lib[0].fv3_cube_link(ctxs, (C.c_int * 6)(1, 2, 3, 4, 5, 6), 6) == 0
    up, dn, keep = lib[0].fv3_upload_state, lib[0].fv3_download_state, []
    up.restype = dn.restype = C.c_int
    for t in range(1, 7):
        st = case.states[t - 1]
        s = abi.State()
        for fld, key in (("u", "u"), ("v", "v"), ("w", "w"), ("delz", "delz"), ("pt", "pt"), ("delp", "delp"), ("phis", "phis")):
            a = np.ascontiguousarray(st[key], dtype=np.float64); keep.append(a)
            setattr(s, fld, _p(a))
        assert up(eng[t].ctx, C.byref(s)) == 0, eng[t].last_error()
    fn = lib[0].fv3_dyn_core
    fn.restype = C.c_int
    assert fn(ctxs, 6, C.c_double(600.0), C.c_int(2), C.c_int(0)) == 0, eng[1].last_error()
    assert fn(ctxs, 6, C.c_double(600.0), C.c_int(2), C.c_int(4)) == -2          # bits 0 (FV3_DYN_GRAPH) and 1 (FV3_DYN_END_STEP) are defined
    oc.dyn_core(600.0, 2)
    assert fn(ctxs, 6, C.c_double(600.0), C.c_int(2), C.c_int(1)) == 0           # second call as a CUDA graph (tests/test_cuda_graph_gpu.py)
    oc.dyn_core(600.0, 2)
    b = case.bounds
    reg = H.regions_state(b)
    for t in range(1, 7):
        e = eng[t]
        out = {f: np.zeros(e.shape(f)) for f in ("U", "V", "W", "DELZ", "PT", "DELP", "MFX", "MFY", "CX", "CY")}
        s = abi.State()
        for fld, f in (("u", "U"), ("v", "V"), ("w", "W"), ("delz", "DELZ"), ("pt", "PT"), ("delp", "DELP"), ("mfx", "MFX"), ("mfy", "MFY"),
                       ("cx", "CX"), ("cy", "CY")):
            setattr(s, fld, _p(out[f]))
        assert dn(e.ctx, C.byref(s)) == 0, e.last_error()
        for f, a in out.items():
            assert _err(_section(e, f, a, reg[f]), _section(oc.eng[t], f, oc.eng[t].get(f), reg[f])) <= (1e-9 if f == "W" else 1e-10), (t, f)
    oc.close()
    for e in eng.values():
        e.close()
