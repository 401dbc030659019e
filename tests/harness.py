"""Parity harness (TEST INFRASTRUCTURE): drives the CUDA product and the CPU oracle through the
same field / stage vocabulary and compares them.  Used by tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs only."""
from __future__ import annotations

import ctypes as C
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gfdl_atmos_cubed_sphere_b200 import abi, grid as G, init_state as I, cubed_sphere as cs  # noqa: E402

from gfdl_atmos_cubed_sphere_b200.cube import Case, CudaCube, FLAGSETS, cube_grid  # noqa: E402,F401  (the product's driver)


def load_oracle(fast=False):
    name = "libfv3_oracle_fast.so" if fast else "libfv3_oracle.so"
    path = os.path.join(ROOT, "oracle", "_build", name)
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j4"])
    return C.CDLL(path), "fv3o_"


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def sub(eng, name, arr, i0, i1, j0, j1):
    """Section (i0:i1, j0:j1) (Fortran, inclusive) of a field array returned by Engine.get."""
    ilo, ni, jlo, nj, nk, kmid = eng.dims(name)
    if kmid:
        return arr[j0 - jlo:j1 - jlo + 1, :, i0 - ilo:i1 - ilo + 1]
    return arr[:, j0 - jlo:j1 - jlo + 1, i0 - ilo:i1 - ilo + 1]


def scaled_err(a, b):
    den = float(np.max(np.abs(b)))
    if not np.isfinite(den) or den == 0.0:
        den = 1.0
    return float(np.max(np.abs(a - b))) / den


def compare(e_ref, e_new, regions):
    """regions: {field: (i0, i1, j0, j1)} -> {field: scaled error}"""
    out = {}
    for f, (i0, i1, j0, j1) in regions.items():
        a = sub(e_new, f, e_new.get(f), i0, i1, j0, j1)
        b = sub(e_ref, f, e_ref.get(f), i0, i1, j0, j1)
        out[f] = scaled_err(a, b)
    log = os.environ.get("FV3_PARITY_LOG")   # per-field error record of a test run (how the tolerances in tests/ were set)
    if log:
        import json
        with open(log, "a") as fh:
            fh.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "err": out}) + "\n")
    return out


def regions_c_sw(b):
    is_, ie, js, je = b["is_"], b["ie"], b["js"], b["je"]
    return {"DELPC": (is_ - 1, ie + 1, js - 1, je + 1), "PTC": (is_ - 1, ie + 1, js - 1, je + 1),
            "OMGA": (is_ - 1, ie + 1, js - 1, je + 1), "UC": (is_ - 1, ie + 2, js - 1, je + 1),
            "VC": (is_ - 1, ie + 1, js - 1, je + 2), "UA": (is_ - 2, ie + 2, js - 2, je + 2),
            "VA": (is_ - 2, ie + 2, js - 2, je + 2), "UT": (is_ - 1, ie + 2, js - 1, je + 1),
            "VT": (is_ - 1, ie + 1, js - 1, je + 2), "DIVGD": (is_, ie + 1, js, je + 1)}


def regions_d_sw(b, use_cond=False, heat=False):
    is_, ie, js, je, jsd, jed, isd, ied = b["is_"], b["ie"], b["js"], b["je"], b["jsd"], b["jed"], b["isd"], b["ied"]
    r = {"DELP": (is_, ie, js, je), "PT": (is_, ie, js, je), "W": (is_, ie, js, je), "U": (is_, ie, js, je + 1),
         "V": (is_, ie + 1, js, je), "CRX": (is_, ie + 1, jsd, jed), "XFX": (is_, ie + 1, jsd, jed),
         "CRY": (isd, ied, js, je + 1), "YFX": (isd, ied, js, je + 1), "MFX": (is_, ie + 1, js, je),
         "MFY": (is_, ie, js, je + 1), "CX": (is_, ie + 1, jsd, jed), "CY": (isd, ied, js, je + 1)}
    if use_cond:
        r["QCON"] = (is_, ie, js, je)
    if heat:
        r["HEAT"] = (is_, ie, js, je)
    return r


def parity_c_sw_d_sw(n=24, npz=8, flagset="A", dt=20.0, tile=1, flags_override=None):
    """c_sw -> p-grad-free hand-over -> d_sw on one face, CUDA vs oracle; returns scaled errors."""
    case = Case(n, npz, flagset, flags_override=flags_override)
    eo = case.engine(load_oracle(), tile)
    eg = case.engine(abi.load_library(), tile)
    res = {}
    for e in (eo, eg):
        case.load_state(e, tile)
        e.call("c_sw", 0.5 * dt)
    res.update({"c_sw." + k: v for k, v in compare(eo, eg, regions_c_sw(case.bounds)).items()})
    # hand the ORACLE's c_sw outputs to both d_sw's so the stages are compared independently
    for f in ("UC", "VC", "UA", "VA", "DIVGD", "UT", "VT", "DELPC", "PTC"):
        eg.put(f, eo.get(f))
    for e in (eo, eg):
        e.call("d_sw", dt)
    use_cond = bool(case.flags.get("use_cond"))
    heat = case.flags.get("d_con", 0.0) > 1e-5
    res.update({"d_sw." + k: v for k, v in compare(eo, eg, regions_d_sw(case.bounds, use_cond, heat)).items()})
    eo.close(); eg.close()
    return res


# ---- python-driven oracle dyn_core (mirrors csrc/dyn_core.cu; reference dyn_core.F90:313-1286) ----
class OracleCube:
    """Faces of the cube on the CPU oracle with the NumPy halo exchange."""

    def __init__(self, case, tiles=(1, 2, 3, 4, 5, 6), fast=False, numpy_halo=False):
        self.case = case
        self.tiles = list(tiles)
        lib = load_oracle(fast)
        self.lib = lib
        self.numpy_halo = numpy_halo or os.environ.get("FV3O_NUMPY_HALO") == "1" or len(self.tiles) != 6
        self.eng = {t: case.engine(lib, t) for t in self.tiles}
        self.ex = cs.Exchanger(case.n, 3) if len(self.tiles) == 6 else None
        for t in self.tiles:
            case.load_state(self.eng[t], t)

    # the exchange runs inside the oracle library (fv3o_halo_exchange) on the tables of cubed_sphere.py; FV3O_NUMPY_HALO=1
    # selects the NumPy gather on the same tables (tests/test_grid_and_halo.py compares the two)
    _table_ids = {}

    def _table(self, pos_x, pos_y=None, kind="scalar", boundary_only=False):
        key = (self.lib[0]._name, self.case.n, pos_x, pos_y, kind, boundary_only)   # the parity and the timing build keep separate registries
        ids = OracleCube._table_ids
        if key not in ids:
            tabs = self.ex.tables(pos_x, pos_y, kind, boundary_only=boundary_only)
            tid = len(ids) + 1
            fn = self.lib[0].fv3o_halo_set_table
            fn.restype = C.c_int
            for t in range(1, 7):
                for ci, tb in enumerate(tabs[t]):
                    dst = np.ascontiguousarray(tb.dst, dtype=np.int64); src = np.ascontiguousarray(tb.src, dtype=np.int64)
                    st = np.ascontiguousarray(tb.src_tile, dtype=np.int32); sc = np.ascontiguousarray(tb.src_comp, dtype=np.int32)
                    sg = np.ascontiguousarray(tb.sign, dtype=np.float64)
                    rc = fn(C.c_int(tid), C.c_int(t), C.c_int(ci), C.c_longlong(dst.size), dst.ctypes.data_as(C.c_void_p),
                            st.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p),
                            sg.ctypes.data_as(C.c_void_p))
                    assert rc == 0
            ids[key] = tid
        return ids[key]

    def _lib_exchange(self, fx, fy, tid):
        fn = self.lib[0].fv3o_halo_exchange
        fn.restype = C.c_int
        ctxs = (C.c_void_p * 6)(*[self.eng[t].ctx for t in (1, 2, 3, 4, 5, 6)])
        rc = fn(ctxs, C.c_int(abi.FIELD_ID[fx]), C.c_int(abi.FIELD_ID[fy] if fy else -1), C.c_int(tid))
        assert rc == 0, f"fv3o_halo_exchange rc={rc}"

    def _exchange_scalar(self, name, pos=cs.CENTER):
        if not self.numpy_halo:
            return self._lib_exchange(name, None, self._table(pos))
        arrs = [self.eng[t].get(name) for t in self.tiles]
        self.ex.scalar(arrs, pos)
        for t, a in zip(self.tiles, arrs):
            self.eng[t].put(name, a)

    def _exchange_pair(self, nx, ny, px, py, boundary_only=False):
        if not self.numpy_halo:
            return self._lib_exchange(nx, ny, self._table(px, py, "vector", boundary_only))
        xs = [self.eng[t].get(nx) for t in self.tiles]
        ys = [self.eng[t].get(ny) for t in self.tiles]
        self.ex.pair(xs, ys, px, py, kind="vector", boundary_only=boundary_only)
        for t, a, b in zip(self.tiles, xs, ys):
            self.eng[t].put(nx, a); self.eng[t].put(ny, b)

    def halo(self, group):
        if self.ex is None:
            return
        f = self.case.flags
        if group == "UVW":
            self._exchange_pair("U", "V", cs.NORTH, cs.EAST)
            if not f["hydrostatic"]:
                self._exchange_scalar("W")
        elif group == "GZ":
            self._exchange_scalar("GZ")
        elif group == "DIVGD_UCVC":
            if f["nord"] > 0:
                self._exchange_scalar("DIVGD", cs.CORNER)
            self._exchange_pair("UC", "VC", cs.EAST, cs.NORTH)
        elif group == "DELP_PT":
            self._exchange_scalar("DELP"); self._exchange_scalar("PT")
            if f.get("use_cond"):
                self._exchange_scalar("QCON")
        elif group == "ZH_PKC":
            self._exchange_scalar("ZH"); self._exchange_scalar("PKC")
        elif group == "UV_EDGE":
            self._exchange_pair("U", "V", cs.NORTH, cs.EAST, boundary_only=True)
        elif group == "HEAT":
            self._exchange_scalar("HEAT")
        elif group == "OMGA":
            self._exchange_scalar("OMGA")

    def all(self, stage, *args):
        for t in self.tiles:
            self.eng[t].call(stage, *args)

    def dyn_core(self, bdt, n_split, timers=None, end_step=False):
        """end_step: the reference's argument of that name (last dyn_core call of the k_split loop): omega diagnostic on the last substep"""
        import time
        dt = bdt / n_split
        dt2 = 0.5 * dt
        F = abi.FIELD_ID

        def run(name, stage, *args):
            t0 = time.perf_counter()
            self.all(stage, *args)
            if timers is not None:
                timers[name] = timers.get(name, 0.0) + time.perf_counter() - t0

        for f in ("MFX", "MFY", "CX", "CY", "HEAT"):
            self.all("zero_field", F[f])
        hydro = bool(self.case.flags["hydrostatic"])
        beta = float(self.case.flags.get("beta", 0.0))
        if beta > 0.0:
            for f in ("DU", "DV"):
                self.all("zero_field", F[f])
        for it in range(1, n_split + 1):
            last = it == n_split
            if it == 1:
                self.halo("DELP_PT")
            omega = last and end_step and self.case.flags.get("sw_test_case") != 1
            omega_new = omega and not self.case.flags.get("use_old_omega", 1)
            if omega:
                self.all("omega_begin")          # dyn_core.F90:409-422: pem from delp before the last substep (use_old_omega = T)
            if self.case.flags.get("sw_test_case") == 1:   # SW_DYNAMICS, test_case 1: d_sw + delp halo only (dyn_core.F90:394, 569, 998)
                run("D_SW", "d_sw", dt)
                self.halo("DELP_PT")
                continue
            self.halo("UVW")
            if hydro:   # dyn_core.F90:478-480, :905-907, :1017-1021
                run("C_SW", "c_sw", dt2)
                run("GEOPK_C", "geopk", 1)
                run("PG_C", "p_grad_c", dt2)
                self.halo("DIVGD_UCVC")
                if self.case.flags.get("d_ext", 0.0) > 0.0:
                    self.all("ext_mode_prepare")        # dyn_core.F90:745-747: delp at the cell corners, before d_sw
                if omega_new:
                    self.all("omega_new", 0, dt)     # use_old_omega = F (dyn_core.F90:735-742): omga = delp before d_sw
                run("D_SW", "d_sw", dt)
                if omega_new:
                    self.all("omega_new", 1, dt)     # :774-781: times the convergence of the area fluxes / dt
                if self.case.flags.get("d_ext", 0.0) > 0.0:
                    self.all("ext_mode_divg2")          # :828-847: mass-weighted vertical mean of the divergence
                self.halo("DELP_PT")
                run("GEOPK_D", "geopk", 0)
                if last:
                    self.all("pk_from_pkc")   # dyn_core.F90:1001-1010
                if beta > 0.0:
                    run("PG_D", "split_p_grad", dt, 0.0 if it == 1 else beta)   # grad1_p_update, dyn_core.F90:1018-1019
                else:
                    run("PG_D", "one_grad_p", dt)
                if last:
                    self.halo("UV_EDGE")
                if omega:
                    self.all("omega_end", dt)    # :1182-1195
                continue
            if it == 1:
                self.all("gz_init")
                self.halo("GZ")
            run("C_SW", "c_sw", dt2)
            if it == 1:
                self.all("copy_field", F["ZH"], F["GZ"])
            else:
                self.all("copy_field", F["GZ"], F["ZH"])
            run("UPDATE_DZ_C", "update_dz_c", dt2)
            run("Riem_Solver_C", "riem_solver_c", dt2)
            run("PG_C", "p_grad_c", dt2)
            self.halo("DIVGD_UCVC")
            if omega_new:
                self.all("omega_new", 0, dt)     # use_old_omega = F (dyn_core.F90:735-742): omga = delp before d_sw
            run("D_SW", "d_sw", dt)
            if omega_new:
                self.all("omega_new", 1, dt)     # :774-781: times the convergence of the area fluxes / dt
            self.halo("DELP_PT")
            run("UPDATE_DZ", "update_dz_d", dt)
            run("Riem_Solver3", "riem_solver3", dt, 1 if last else 0)
            self.halo("ZH_PKC")
            if last:
                self.all("pe_halo")
            self.all("pk3_halo")
            self.all("gz_from_zh")
            if beta > 0.0:
                run("PG_D", "split_p_grad", dt, 0.0 if it == 1 else beta)       # dyn_core.F90:1027-1028
            else:
                run("PG_D", "nh_p_grad", dt)
            if last:
                self.halo("UV_EDGE")
            if omega:
                self.all("omega_end", dt)        # :1182-1195
        self.dcon_heating(bdt)

    def n_con(self):
        """dyn_core.F90:296-307"""
        f = self.case.flags
        if f.get("convert_ke") or (f.get("do_vort_damp") and f.get("vtdm4", 0.0) > 1e-4):
            return self.case.npz
        if f.get("d2_bg_k1", 0.0) < 1e-3:
            return 0
        return 1 if f.get("d2_bg_k2", 0.0) < 1e-3 else 2

    def del2_cubed(self, field, cd, nmax):
        """del2_cubed incl. its halo update (dyn_core.F90:2356-2465) on every face."""
        self.halo(field if field != "HEAT" else "HEAT")
        self.all("del2_cubed", abi.FIELD_ID[field], float(cd), int(nmax))

    def dcon_heating(self, bdt):
        """dyn_core.F90:1300-1356 (the SW_DYNAMICS branch returns before it)."""
        f = self.case.flags
        if f.get("sw_test_case") == 1 or not (f.get("d_con", 0.0) > 1e-5) or self.n_con() == 0:
            return
        da_min = self.case.tiles[0].da_min
        self.del2_cubed("HEAT", 0.20 * da_min, min(3, f["nord"] + 1))
        self.all("dcon_heating", float(bdt))

    def fv_dynamics(self, bdt, k_split, n_split, kord_mt, kord_wz, kord_tm, kord_tr, hord_tr, nf_omega, sphum=-1, zvir=0.0):
        """fv_dynamics.F90:303-398 + :445-662 on the oracle side, the mirror of fv3_fv_dynamics[_qv] (no condensates)."""
        F = abi.FIELD_ID
        if sphum >= 0:
            self.select_tracer(sphum)
            self.all("pt_to_theta", float(zvir))
            self.select_tracer(0)
        else:
            self.all("pt_to_theta", 0.0)
        mdt = bdt / k_split
        for n_map in range(1, k_split + 1):
            last = int(n_map == k_split)
            self.all("copy_field", F["DP1"], F["DELP"])
            self.dyn_core(mdt, n_split, end_step=bool(last))
            if hord_tr:
                self.tracer_2d(hord_tr)
            self.all("lagrangian_to_eulerian_qv", last, kord_mt, kord_wz, kord_tm, getattr(self, "nq", 1) if hord_tr else 0, kord_tr,
                     int(sphum), float(zvir))
            if last and nf_omega > 0:
                self.del2_cubed("OMGA", 0.18 * self.case.tiles[0].da_min, nf_omega)

    def set_num_tracers(self, nq):
        self.nq = nq
        self.all("set_num_tracers", nq)

    def select_tracer(self, iq):
        self.all("select_tracer", iq)

    def set_tracer_fill(self, on):
        self.all("set_tracer_fill", int(bool(on)))

    def tracer_2d(self, hord):
        """tracer_2d_1L (model/fv_tracer2d.F90:49-295; all tracers of the engines, trdm = 0, id_divg_mean = 0) on the oracle side: the pointwise
        statements in NumPy with the reference's operation order, the fluxes from the oracle's fv_tp_2d, the q halo updates
        from the NumPy exchange, the CFL maximum over the six faces.  Operates on WORK_Q, DP1, CX, CY, MFX, MFY (XFX, YFX
        are overwritten) exactly like fv3_tracer_2d; returns cmax(npz)."""
        case = self.case
        n, npz, b = case.n, case.npz, case.bounds
        is_, ie, js, je, isd, jsd = b["is_"], b["ie"], b["js"], b["je"], b["isd"], b["jsd"]
        I = np.arange(is_, ie + 2); J = np.arange(js, je + 2)
        st = {}
        cmax = np.zeros(npz)
        for t in self.tiles:
            e, g = self.eng[t], case.tiles[t - 1].arr
            cx, cy = e.get("CX"), e.get("CY")          # (npz, jsd:jed, is:ie+1), (npz, js:je+1, isd:ied)
            dxa, dya, dx, dy, sg = g["dxa"], g["dya"], g["dx"], g["dy"], g["sin_sg"]
            xfx = np.where(cx > 0., cx * dxa[:, I - 1 - isd] * dy[:, I - isd] * sg[2][:, I - 1 - isd],
                           cx * dxa[:, I - isd] * dy[:, I - isd] * sg[0][:, I - isd])            # :112-118
            yfx = np.where(cy > 0., cy * dya[J - 1 - jsd, :] * dx[J - jsd, :] * sg[3][J - 1 - jsd, :],
                           cy * dya[J - jsd, :] * dx[J - jsd, :] * sg[1][J - jsd, :])            # :121-127
            acx = np.abs(cx[:, js - jsd:je - jsd + 1, 0:n]); acy = np.abs(cy[:, 0:n, is_ - isd:ie - isd + 1])
            sg5 = sg[4][js - jsd:je - jsd + 1, is_ - isd:ie - isd + 1]
            for k in range(npz):                                                                 # :131-145
                m = np.maximum(acx[k], acy[k])
                if not ((k + 1) < npz // 6):
                    m = m + 1. - sg5
                cmax[k] = max(cmax[k], float(m.max()))
            st[t] = dict(cx=cx, cy=cy, xfx=xfx, yfx=yfx, mfx=e.get("MFX"), mfy=e.get("MFY"), dp1=e.get("DP1"),
                         area=g["area"], rarea=g["rarea"][js - jsd:je - jsd + 1, is_ - isd:ie - isd + 1])
        nsplt = (1. + cmax).astype(np.int64)                                                     # mp_reduce_max :161, :170
        for t in self.tiles:
            s = st[t]
            for k in range(npz):
                if nsplt[k] > 1:                                                                 # :171-190
                    fr = 1. / float(nsplt[k])
                    for nm in ("cx", "xfx", "cy", "yfx", "mfx", "mfy"):
                        s[nm][k] = s[nm][k] * fr
            s["ra_x"] = s["area"][None, :, is_ - isd:ie - isd + 1] + s["xfx"][:, :, :-1] - s["xfx"][:, :, 1:]     # :199-201
            s["ra_y"] = s["area"][None, js - jsd:je - jsd + 1, :] + s["yfx"][:, :-1, :] - s["yfx"][:, 1:, :]     # :203-205
            e = self.eng[t]
            for f, nm in (("CRX", "cx"), ("CRY", "cy"), ("XFX", "xfx"), ("YFX", "yfx"), ("WORK_RAX", "ra_x"), ("WORK_RAY", "ra_y"),
                          ("MFX", "mfx"), ("MFY", "mfy"), ("CX", "cx"), ("CY", "cy")):
                e.put(f, s[nm])
        sl = (slice(None), slice(js - jsd, je - jsd + 1), slice(is_ - isd, ie - isd + 1))
        nq = getattr(self, "nq", 1)
        for it in range(1, int(nsplt.max()) + 1):
            for iq in range(nq):                                                                 # every tracer sees the same dp1 -> dp2 (:206-266)
                if nq > 1:
                    self.select_tracer(iq)
                self._exchange_scalar("WORK_Q")                                                  # q_pack :188 / qn2 :282
                for t in self.tiles:
                    e, s = self.eng[t], st[t]
                    e.call("fv_tp_2d", npz, hord, 1, 0, 0, 0.0)
                    fx, fy = e.get("WORK_FX"), e.get("WORK_FY")
                    q = e.get("WORK_Q")
                    qi, d1 = q[sl], s["dp1"][sl]
                    dp2 = d1 + (s["mfx"][:, :, :-1] - s["mfx"][:, :, 1:] + s["mfy"][:, :-1, :] - s["mfy"][:, 1:, :]) * s["rarea"][None]   # :212
                    qn = (qi * d1 + (fx[:, :, :-1] - fx[:, :, 1:] + fy[:, :-1, :] - fy[:, 1:, :]) * s["rarea"][None]) / dp2          # :232,239,250
                    act = (it <= nsplt)[:, None, None]
                    q[sl] = np.where(act, qn, qi)
                    if iq == nq - 1:
                        s["dp1"][sl] = np.where((it < nsplt)[:, None, None], dp2, d1)            # :268-274, after the last tracer
                    e.put("WORK_Q", q)
        if nq > 1:
            self.select_tracer(0)
        for t in self.tiles:
            self.eng[t].put("DP1", st[t]["dp1"])
        return cmax

    def close(self):
        for e in self.eng.values():
            e.close()


def regions_state(b):
    is_, ie, js, je = b["is_"], b["ie"], b["js"], b["je"]
    return {"DELP": (is_, ie, js, je), "PT": (is_, ie, js, je), "W": (is_, ie, js, je), "U": (is_, ie, js, je + 1),
            "V": (is_, ie + 1, js, je), "DELZ": (is_, ie, js, je), "MFX": (is_, ie + 1, js, je), "MFY": (is_, ie, js, je + 1),
            "CX": (is_, ie + 1, b["jsd"], b["jed"]), "CY": (b["isd"], b["ied"], js, je + 1), "ZH": (is_, ie, js, je)}
    # pkc, pk3, gz are NOT compared after nh_p_grad: the reference interpolates them to cell corners
    # in place (a2b_ord4 replace=.true., dyn_core.F90:1743-1746) and rebuilds them next substep;
    # the CUDA path keeps the B-grid copies in scratch planes instead.
