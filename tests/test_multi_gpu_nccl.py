"""-m gpu, needs >= 2 (>= 6) GPUs on the box: the faces of the cube spread over N ranks (one process per GPU, the library's NCCL
send/recv halo exchange, ncclAllReduce(max) in tracer_2d) give the single-process result bit for bit -- tests/nccl_check.py
under torchrun.  N = 2 (3 + 3 faces: on-rank gathers and off-rank messages mixed) and N = 6 (one face per GPU, every contact
off-rank: BASELINE.json configs[2]/[3] placement), flag-sets A and B, with the peer-mapped exchange and with NCCL send / recv on
the data path, once overlapped, and once with a caller-owned communicator (fv3_comm_attach).  Skipped on boxes with fewer GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(n, flagset, attach=False, p2p=1, overlap=False):
    env = dict(os.environ, FV3_CHECK_FLAGSET=flagset, FV3_CHECK_ATTACH="1" if attach else "0", FV3_HALO_P2P=str(p2p))
    if not attach:
        env["FV3_CHECK_EXPECT_P2P"] = str(p2p)     # the run must really have used the exchange it claims
    if overlap:
        env["FV3_HALO_OVERLAP"] = "1"
    else:
        env.pop("FV3_HALO_OVERLAP", None)
    port = 29700 + (os.getpid() + 7 * n + (3 if attach else 0) + (1 if flagset == "B" else 0) + 11 * p2p + (5 if overlap else 0)) % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "== single-process result: True" in p.stdout, p.stdout[-2000:]


@pytest.mark.parametrize("p2p", [1, 0])
@pytest.mark.parametrize("flagset", ["A", "B"])
@pytest.mark.parametrize("n", [2, 6])
def test_multi_rank_run_is_bit_identical_to_the_single_process_run(n, flagset, p2p):
    """p2p = 1: peer-mapped exchange (CUDA IPC arenas, the default); 0: NCCL send / recv (FV3_HALO_P2P=0)."""
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    _run(n, flagset, p2p=p2p)


def test_overlapped_peer_mapped_exchange_is_bit_identical():
    """The delp / pt exchange on the side stream underneath update_dz_d + Riem_Solver3 (FV3_HALO_OVERLAP=1)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "A", p2p=1, overlap=True)


def test_caller_owned_communicator_through_fv3_comm_attach():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "A", attach=True)
