"""CPU: the flux arithmetic the CUDA tile kernels execute (ppm.cuh, tp_tile.cuh::line_flux_na -- compiled __host__ __device__) run on
the HOST and compared with the oracle's xppm / yppm for every scheme of tp_valid_schemes.  GPU time is scarce; this keeps the
kernels' scalar math pinned to the oracle on every CPU run.  (The device build is unaffected: its SASS is byte-identical with
and without the host annotations.)"""
import ctypes as C
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ALL = [-5, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]
COMMON = [-5, 5, 6, 8, 10]


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("hostppm") / "host_ppm_test")
    subprocess.check_call([NVCC, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out,
                           os.path.join(ROOT, "tests", "host_ppm_test.cu")], stderr=subprocess.DEVNULL)
    return out


def _run(exe, mode, n, iord, p0, p1, *arrays):
    blob = struct.pack("5i", mode, n, iord, p0, p1) + b"".join(np.ascontiguousarray(a, dtype=np.float64).tobytes() for a in arrays)
    r = subprocess.run([exe], input=blob, capture_output=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    return np.frombuffer(r.stdout, dtype=np.float64)


def _host_flux(exe, q, c, iord, stride, rare):
    return _run(exe, 0, q.size, iord, stride, rare, q, c)


def _oracle_flux(q, c, iord, ydir):
    lib, _ = H.load_oracle()
    dp = C.POINTER(C.c_double)
    flux = np.zeros(q.size + 1)
    assert lib.fv3o_ppm_periodic(q.size, q.ctypes.data_as(dp), c.ctypes.data_as(dp), iord, ydir, flux.ctypes.data_as(dp)) == 0
    return flux


def _data(seed, positive):
    rng = np.random.default_rng(seed)
    n = 96
    x = (np.arange(n) + 0.5) / n
    q = np.sin(2 * np.pi * x) + 0.5 * np.sign(np.sin(6 * np.pi * x)) + 0.1 * rng.standard_normal(n)
    if positive:
        q = np.maximum(q, 0.0)           # zeros and positive bumps: the positive-definite constraints act
    c = rng.uniform(-0.9, 0.9, n + 1)    # both upwind directions, face by face
    return np.ascontiguousarray(q), np.ascontiguousarray(c)


@pytest.mark.parametrize("iord", ALL)
def test_kernel_flux_arithmetic_on_the_host_matches_the_oracle(built, exe, iord):
    for seed, positive in ((1, False), (2, True)):
        q, c = _data(seed, positive)
        for stride, ydir in ((1, 0), (7, 1)):
            got = _host_flux(exe, q, c, iord, stride, 1)
            want = _oracle_flux(q, c, iord, ydir)
            err = np.abs(got - want).max() / max(1.0, np.abs(want).max())
            assert err < 1e-14, (iord, seed, stride, err)


@pytest.mark.parametrize("iord", COMMON)
def test_hot_instantiation_equals_general_one(built, exe, iord):
    """RARE = false (what the FAM 0 / 1 kernels compile) gives bit-identical fluxes for the schemes it serves."""
    q, c = _data(3, iord == -5)
    a = _host_flux(exe, q, c, iord, 1, 0)
    b = _host_flux(exe, q, c, iord, 1, 1)
    assert np.array_equal(a, b)


def _cube_line(seed, positive, n=40):
    """A face line with its 3-cell halos (index -2..n+3), a smoothly varying metric, Courant numbers of both signs."""
    rng = np.random.default_rng(seed)
    x = (np.arange(-2, n + 4) - 0.5) / n
    q = np.sin(2 * np.pi * x) + 0.5 * np.sign(np.sin(5 * np.pi * x)) + 0.1 * rng.standard_normal(n + 6)
    if positive:
        q = np.maximum(q, 0.0)
    c = rng.uniform(-0.9, 0.9, n + 1)
    dxa = 1.0e5 * (1.0 + 0.3 * np.cos(np.pi * (x - 0.5))) * (1.0 + 0.02 * rng.standard_normal(n + 6))
    return q, c, dxa


@pytest.mark.parametrize("iord", ALL)
def test_cube_edge_operator_on_the_host_matches_the_oracle(built, exe, iord):
    """ppm::flux_scalar -- what edge_flux runs for the cube-edge faces of the frame tiles -- over a whole face line: the
    one-sided edge formulas at both ends (tp_core.F90:376-392, 636-681) and the interior formulas in between."""
    lib, _ = H.load_oracle()
    dp = C.POINTER(C.c_double)
    for seed, positive in ((11, False), (12, True)):
        q, c, dxa = _cube_line(seed, positive)
        n = c.size - 1
        for stride, ydir in ((1, 0), (5, 1)):
            got = _run(exe, 1, n, iord, stride, 1, q, c, dxa)
            want = np.zeros(n + 1)
            assert lib.fv3o_ppm_cube_line(n, q.ctypes.data_as(dp), c.ctypes.data_as(dp), dxa.ctypes.data_as(dp), iord, ydir,
                                          want.ctypes.data_as(dp)) == 0
            err = np.abs(got - want).max() / max(1.0, np.abs(want).max())
            assert err < 1e-13, (iord, seed, stride, err, int(np.abs(got - want).argmax()))
        if iord in COMMON:
            assert np.array_equal(_run(exe, 1, n, iord, 1, 0, q, c, dxa), _run(exe, 1, n, iord, 1, 1, q, c, dxa))


WIND = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]


@pytest.mark.parametrize("iord", WIND)
@pytest.mark.parametrize("edge_line", [0, 1])
def test_wind_operator_on_the_host_matches_the_oracle(built, exe, iord, edge_line):
    """ppm::flux_wind (xtp_u / ytp_v as k_dsw_ke evaluates them near the cube edges) over a whole face line, on an ordinary
    row and on a face-edge row (j = 1: bl = br = 0 at the corner cells, sw_core.F90:2206-2210); every scheme of xtp_u
    (sw_core.F90:2187-2508), against the oracle's xtp_u AND ytp_v (the kernels use one routine for both directions).
    Then the interior fast path (flux_wind_fast_g) on the faces where k_dsw_ke takes it."""
    lib, _ = H.load_oracle()
    dp = C.POINTER(C.c_double)
    u, c, dx = _cube_line(21 + iord, False)
    n = c.size - 1
    c = c * 0.5 * dx[2:n + 3]            # xtp_u's c is a distance: cfl = c * rdx(upwind)
    rdx = 1.0 / dx
    line = 1 if edge_line else 7
    got = _run(exe, 2, n, iord, edge_line, 0, u, c, dx, rdx)
    for fn in (lib.fv3o_xtp_u_line, lib.fv3o_ytp_v_line):
        want = np.zeros(n + 1)
        assert fn(n, line, u.ctypes.data_as(dp), c.ctypes.data_as(dp), dx.ctypes.data_as(dp), rdx.ctypes.data_as(dp), iord,
                  want.ctypes.data_as(dp)) == 0
        err = np.abs(got - want).max() / max(1.0, np.abs(want).max())
        assert err < 1e-13, (iord, edge_line, fn.__name__, err, int(np.abs(got - want).argmax()))
    if not edge_line:
        fast = _run(exe, 3, n, iord, 0, 0, u, c, dx, rdx)
        sl = slice(3, n - 2)             # faces 4..n-2
        assert np.abs(fast[sl] - want[sl]).max() / max(1.0, np.abs(want).max()) < 1e-13, iord
