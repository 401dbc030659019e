"""CPU: the vertical-remap column operators the CUDA kernels execute (csrc/remap_col.cuh, compiled __host__ __device__) run on the
HOST and compared with the oracle (oracle/remap.cpp) for every scheme 3..15 (ppm_profile for <= 7, the cs / scalar profiles above) and every boundary mode -- BIT FOR BIT: remap.cu is
built without FMA contraction precisely so that its decisions at the tie-prone switches of schemes 11 / 12 are the oracle's
(tests/test_remap_gpu.py), and the host build has no FMA either.  GPU time is scarce; this keeps the kernels' column arithmetic
pinned on every CPU run and lets the sweeps be restructured (register carries, batching) without a device."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import harness as H
import test_remap_oracle as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("hostremap") / "host_remap_test")
    subprocess.check_call([NVCC, "-std=c++17", "-O1", "--fmad=false", "-Xcompiler", "-ffp-contract=off", "-gencode",
                           "arch=compute_100a,code=sm_100a", "-o", out, os.path.join(ROOT, "tests", "host_remap_test.cu")],
                          stderr=subprocess.DEVNULL)
    return out


@pytest.fixture(scope="module")
def state():
    case, oc = T._cube(substeps=2)
    yield case, oc
    oc.close()


@pytest.mark.parametrize("kord", T.KORDS)
@pytest.mark.parametrize("mode,iv", [(0, 1), (1, 1), (1, -1), (1, -2), (2, 0), (0, 0), (1, 2)])
def test_host_execution_of_the_cuda_remap_matches_the_oracle_bit_for_bit(exe, state, kord, mode, iv):
    case, oc = state
    e = oc.eng[4]
    n, km = T.N, T.NPZ
    pe = T._pe(e)                                   # (km + 1, n, n)
    p2 = T._hybrid(case, pe)
    rng = np.random.default_rng(100 * kord + 10 * mode + iv + 5)
    q = np.abs(H.sub(e, "PT", e.get("PT"), 1, n, 1, n)) * rng.uniform(0.0, 1.0, (km, n, n)) ** 3      # sharp, positive
    T._set_q(e, q)
    ws = np.full((n, n), 0.2)
    if iv == -2:
        full = e.get("WS"); H.sub(e, "WS", full, 1, n, 1, n)[...] = ws[None]; e.put("WS", full)
    qmin = 1.0 if mode == 0 else 0.0
    e.call("remap_work_q", mode, iv, kord, qmin)
    want = T._sec(e, "WORK_Q")
    blob = struct.pack("6i", km, n * n, iv, kord, int(mode != 1), 0) + struct.pack("d", qmin) + \
        pe.reshape(km + 1, -1).tobytes() + p2.reshape(km + 1, -1).tobytes() + np.ascontiguousarray(q).reshape(km, -1).tobytes() + \
        (ws if iv == -2 else np.zeros((n, n))).reshape(-1).tobytes()
    r = subprocess.run([exe], input=blob, capture_output=True)
    assert r.returncode == 0, r.stderr
    got = np.frombuffer(r.stdout, dtype=np.float64).reshape(km, n, n)
    assert np.array_equal(got, want), float(np.abs(got - want).max())


def test_host_execution_of_the_cuda_fillz_matches_the_oracle_bit_for_bit(exe, state):
    """fillz_column (csrc/remap_col.cuh, what k_fillz runs per column) on the host against oracle/remap.cpp::fillz_column: columns
    with negative layers next to empty and full ones, a column with nothing to borrow, non-uniform thicknesses"""
    case, oc = state
    e = oc.eng[6]
    n, km = T.N, T.NPZ
    rng = np.random.default_rng(77)
    q0 = e.get("WORK_Q")
    dp = T._sec(e, "DELP")
    q = rng.uniform(-0.4, 1.0, (km, n, n)) * rng.integers(0, 2, (km, n, n))
    q[:, 0, 0] = -np.abs(q[:, 0, 0])
    q[:, 1, :] = np.abs(q[:, 1, :])                  # a row that needs nothing
    q[-1, 2, :] = -0.2; q[-2, 2, :] = 0.5            # the bottom-layer branch
    T._set_q(e, q)
    e.call("fillz")
    want = T._sec(e, "WORK_Q")
    e.put("WORK_Q", q0)
    assert (want != q).any()
    blob = struct.pack("6i", km, n * n, 99, 0, 0, 0) + struct.pack("d", 0.0) + np.zeros((km + 1, n * n)).tobytes() + \
        np.concatenate([dp.reshape(km, -1), np.zeros((1, n * n))]).tobytes() + np.ascontiguousarray(q).reshape(km, -1).tobytes() + \
        np.zeros(n * n).tobytes()
    r = subprocess.run([exe], input=blob, capture_output=True)
    assert r.returncode == 0, r.stderr
    got = np.frombuffer(r.stdout, dtype=np.float64).reshape(km, n, n)
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    assert np.array_equal(np.signbit(got), np.signbit(want))
