"""ctypes mirror of include/fv3_dyncore.h and a thin engine wrapper.

``Engine`` drives a library that exports the field/stage vocabulary of the header under a
symbol prefix.  The package only ever loads the CUDA library (prefix ``fv3_``); the test
harness reuses the same class for its CPU checker.  The product path fails loudly when the
CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_dp = C.POINTER(C.c_double)

FIELDS = ["U", "V", "W", "DELZ", "PT", "DELP", "QCON", "CAPPA", "PHIS", "OMGA", "UA", "VA", "UC", "VC",
          "MFX", "MFY", "CX", "CY", "DELPC", "PTC", "UT", "VT", "DIVGD", "CRX", "CRY", "XFX", "YFX",
          "GZ", "ZH", "PKC", "PK3", "WS3", "WS", "PE", "PELN", "PK", "PKZ", "HEAT", "DISS",
          "WORK_Q", "WORK_FX", "WORK_FY", "WORK_RAX", "WORK_RAY", "DP1", "DU", "DV"]
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}

HALO_GROUPS = ["UVW", "GZ", "DIVGD_UCVC", "DELP_PT", "ZH_PKC", "UV_EDGE", "TRACER", "HEAT", "OMGA"]
HALO_ID = {n: i for i, n in enumerate(HALO_GROUPS)}


class Bounds(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("npx", "npy", "npz", "ng", "is_", "ie", "js", "je", "isd", "ied", "jsd", "jed", "grid_type",
                 "bounded_domain", "sw_corner", "se_corner", "ne_corner", "nw_corner", "stretched_grid", "tile")]


_GRID_PTRS = ["area", "rarea", "dxa", "dya", "rdxa", "rdya", "cosa_s", "rsin2", "f0", "sin_sg", "cos_sg",
              "dy", "rdy", "dxc", "rdxc", "cosa_u", "sina_u", "rsin_u", "divg_v", "del6_v",
              "dx", "rdx", "dyc", "rdyc", "cosa_v", "sina_v", "rsin_v", "divg_u", "del6_u",
              "area_c", "rarea_c", "fC", "cosa", "sina", "rsina",
              "edge_w", "edge_e", "edge_s", "edge_n", "grid", "agrid", "ec1", "ec2", "en1", "en2"]


class Grid(C.Structure):
    _fields_ = [(n, _dp) for n in _GRID_PTRS] + [("da_min", C.c_double), ("da_min_c", C.c_double)]


_FLAG_INTS = ["hord_mt", "hord_vt", "hord_tm", "hord_dp", "hord_tr", "nord", "n_sponge", "m_split",
              "hydrostatic", "do_vort_damp", "use_cond", "moist_kappa", "inline_q", "do_f3d",
              "use_logp", "convert_ke", "prevent_diss_cooling", "do_diss_est", "is_ideal_case",
              "use_old_omega", "fill_dp", "sw_test_case"]
_FLAG_DBLS = ["d4_bg", "d2_bg", "dddmp", "d2_bg_k1", "d2_bg_k2", "vtdm4", "d_con", "ke_bg", "d_ext",
              "a_imp", "p_fac", "beta", "lim_fac", "fast_tau_w_sec", "rf_cutoff", "d2bg_zq", "delt_max",
              "rdgas", "cp_air", "grav", "kappa", "radius", "omega", "pi", "ptop"]


class Flags(C.Structure):
    _fields_ = [(n, C.c_int) for n in _FLAG_INTS] + [(n, C.c_double) for n in _FLAG_DBLS] + \
               [("ak", _dp), ("bk", _dp)]


_STATE_PTRS = ["u", "v", "w", "delz", "pt", "delp", "q_con", "cappa", "phis", "omga", "ua", "va", "uc", "vc",
               "mfx", "mfy", "cx", "cy", "pe", "peln", "pk", "pkz", "ws", "heat_source", "diss_est"]


class State(C.Structure):
    _fields_ = [(n, _dp) for n in _STATE_PTRS]


# Flag-set A of SURVEY 8(d) ("traditional climate / monotonic")
FLAGSET_A = dict(hord_mt=10, hord_vt=10, hord_tm=10, hord_dp=10, hord_tr=8, nord=2, n_sponge=0, m_split=0,
                 hydrostatic=0, do_vort_damp=0, use_cond=0, moist_kappa=0, inline_q=0, do_f3d=0, use_logp=0,
                 convert_ke=0, prevent_diss_cooling=0, do_diss_est=0, is_ideal_case=0, use_old_omega=1, fill_dp=0,
                 d4_bg=0.12, d2_bg=0.0, dddmp=0.0, d2_bg_k1=0.20, d2_bg_k2=0.015, vtdm4=0.0, d_con=0.0, ke_bg=0.0,
                 d_ext=0.0, a_imp=1.0, p_fac=0.05, beta=0.0, lim_fac=1.0, fast_tau_w_sec=0.0, rf_cutoff=3.0e3,
                 d2bg_zq=0.0, delt_max=1.0)
# Flag-set B ("effectively inviscid": exercises the damping / dissipative-heating paths)
FLAGSET_B = dict(FLAGSET_A, hord_mt=5, hord_vt=5, hord_tm=5, hord_dp=-5, hord_tr=-5, nord=3, d4_bg=0.15,
                 do_vort_damp=1, vtdm4=0.03, d_con=1.0, prevent_diss_cooling=1, dddmp=0.2, a_imp=0.75)


def _ptr(a):
    return a.ctypes.data_as(_dp)


def load_library():
    """The CUDA product library.  Fails loudly when it is missing: no CPU fallback."""
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.environ.get("FV3_B200_LIB") or os.path.join(here, "csrc", "libfv3_b200.so")   # override: kernel-variant A/B runs
    if not os.path.exists(path):
        raise RuntimeError(
            f"CUDA library {path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the product path.")
    return C.CDLL(path), "fv3_"


class Engine:
    """One face (tile) of the cube on one library (CUDA product or CPU oracle)."""

    def __init__(self, lib, prefix, bounds: dict, tile_grid, flags: dict, npz: int, ak, bk, ptop,
                 consts: dict, tile: int = 1, device: int = 0):
        self.lib, self.prefix = lib, prefix
        b = Bounds()
        for k, v in bounds.items():
            setattr(b, k, int(v))
        b.npz = npz
        b.bounded_domain = 0
        cube = 1 if bounds["grid_type"] < 3 else 0
        b.sw_corner = b.se_corner = b.ne_corner = b.nw_corner = cube
        b.stretched_grid = 0
        b.tile = tile
        self.bounds = b
        g = Grid()
        self._keep = []
        for n in _GRID_PTRS:
            arr = np.ascontiguousarray(tile_grid.arr[n], dtype=np.float64)
            self._keep.append(arr)
            setattr(g, n, _ptr(arr))
        g.da_min, g.da_min_c = float(tile_grid.da_min), float(tile_grid.da_min_c)
        self.grid = g
        f = Flags()
        for n in _FLAG_INTS:
            setattr(f, n, int(flags.get(n, 0)))
        for n in _FLAG_DBLS:
            if n in flags:
                setattr(f, n, float(flags[n]))
        for n in ("rdgas", "cp_air", "grav", "kappa", "radius", "omega", "pi"):
            setattr(f, n, float(consts[n]))
        f.ptop = float(ptop)
        self._ak = np.ascontiguousarray(ak, dtype=np.float64)
        self._bk = np.ascontiguousarray(bk, dtype=np.float64)
        f.ak, f.bk = _ptr(self._ak), _ptr(self._bk)
        self.flags = f
        self.ctx = C.c_void_p()
        fn = getattr(lib, prefix + "create")
        fn.restype = C.c_int
        rc = fn(C.byref(b), C.byref(g), C.byref(f), C.c_int(device), C.byref(self.ctx))
        if rc != 0:
            raise RuntimeError(f"{prefix}create failed rc={rc}")
        self._dims = {}

    # -- plumbing
    def _fn(self, name, restype=C.c_int):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        return fn

    def last_error(self):
        fn = self._fn("last_error", C.c_char_p)
        return fn(self.ctx).decode()

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{self.prefix}{what} failed rc={rc}: {self.last_error()}")

    def dims(self, name):
        if name not in self._dims:
            d = (C.c_int * 6)()
            self.check(self._fn("field_dims")(self.ctx, FIELD_ID[name], d), "field_dims")
            self._dims[name] = tuple(d)
        return self._dims[name]

    def shape(self, name):
        ilo, ni, jlo, nj, nk, kmid = self.dims(name)
        return (nj, nk, ni) if kmid else (nk, nj, ni)

    def put(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        shp = self.shape(name)
        if a.size != int(np.prod(shp)):
            raise ValueError(f"field {name}: expected shape {shp}, got {a.shape}")
        self.check(self._fn("put_field")(self.ctx, FIELD_ID[name], _ptr(a)), "put_field")

    def get(self, name):
        out = np.empty(self.shape(name), dtype=np.float64)
        self.check(self._fn("get_field")(self.ctx, FIELD_ID[name], _ptr(out)), "get_field")
        return out

    def call(self, stage, *args):
        cargs = [self.ctx]
        for a in args:
            if isinstance(a, float):
                cargs.append(C.c_double(a))
            elif isinstance(a, (int, np.integer, bool)):
                cargs.append(C.c_int(int(a)))
            else:
                cargs.append(a)
        self.check(self._fn(stage)(*cargs), stage)

    def sync(self):
        self.check(self._fn("sync")(self.ctx), "sync")

    def close(self):
        if self.ctx:
            self._fn("destroy", None)(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
