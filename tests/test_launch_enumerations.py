"""CPU: the compact launch enumerations of the CUDA library (frame tile grids, dense frame points, interior / frame
transport tile maps) cover every tile / point exactly once.  The check is host code of the library headers, compiled with
nvcc and run here (no GPU involved)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_frame_and_tile_enumerations(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "host_enum_test")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "host_enum_test.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
