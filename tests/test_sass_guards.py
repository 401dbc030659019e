"""CPU: static guards on the compiled hot kernels (cuobjdump -sass on the in-tree library; nvcc cross-compiles here).
They pin two things measured on the B200: the tile kernels stage their inputs with cp.async (LDGSTS), and the hot
instantiations -- schemes 8/10 (FAM = 1) and 5/6/-5 (FAM = 0), interior tiles -- carry no local-memory traffic and none of the
out-of-line code of the less common schemes (with it d_sw ran 4.6 % slower: profiles/r1_dsw_ncu_summary.md)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gfdl_atmos_cubed_sphere_b200", "csrc", "libfv3_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _sass(mangled):
    if not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)):
        pytest.skip("library or cuobjdump not available")
    out = subprocess.run([CUOBJDUMP, "-sass", "-fun", mangled, LIB], capture_output=True, text=True).stdout
    ins = [re.sub(r"/\*[0-9a-fx]*\*/", "", l).strip() for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    assert len(ins) > 500, f"{mangled}: not found in the library"
    return ins


HOT = ["_Z15k_dsw_transportILi1ELb0EEv3Lay7DevGridN3tpt7TileMapE5DswTr",     # delp / w / pt transport, monotone, interior tiles
       "_Z15k_dsw_transportILi0ELb0EEv3Lay7DevGridN3tpt7TileMapE5DswTr",     # same, unlimited family (flag-set B)
       "_Z13k_dsw_vort_uvILi1ELb0EEv3Lay7DevGridN3tpt7TileMapEPKdS5_S5_S5_S5_S5_S5_S5_PdS6_i"]


@pytest.mark.parametrize("kernel", HOT)
def test_hot_tile_kernels_stage_with_cp_async_and_stay_in_registers(kernel):
    ins = _sass(kernel)
    assert any(i.startswith("LDGSTS") for i in ins), "cp.async staging (LDGSTS) is gone"
    local = [i for i in ins if re.match(r"(LDL|STL)\b", i)]
    # the unlimited-family instance has one 4-byte spill (1 STL + 1 LDL) since before; the out-of-line scheme code costs dozens
    assert len(local) <= 2, f"local-memory traffic in a hot instantiation: {local[:5]}"
    # fp64 arithmetic is what the kernel is for; a build without DFMA/DADD would mean the wrong instantiation
    assert sum(i.startswith(("DFMA", "DADD", "DMUL")) for i in ins) > 200


def test_general_instantiation_exists_and_is_separate():
    """FAM = 2 carries the less common schemes out of line (calls with by-reference results => it does use the stack)."""
    ins = _sass("_Z15k_dsw_transportILi2ELb0EEv3Lay7DevGridN3tpt7TileMapE5DswTr")
    assert any(re.match(r"(LDL|STL)\b", i) for i in ins)


def test_remap_kernels_keep_their_occupancy_budget():
    """The remap kernels are bound by occupancy x memory-level parallelism (csrc/remap.cu header): the in-step kernels are held to 64
    registers (launch bounds: 8 CTAs of 128 threads per SM) and a stack of at most 128 bytes -- a tracer loop inside the kernel once
    cost 480 bytes of stack and 20 % of the remap time (profiles/r2_dsw_summary.md).  Both sets of instantiations (cs / scalar
    profiles; with ppm_profile) and k_fillz are checked."""
    if not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)):
        pytest.skip("library or cuobjdump not available")
    out = subprocess.run([CUOBJDUMP, "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    step = {k: v for k, v in usage.items() if re.search(r"k_remap_(cells|wind)", k)}
    assert len(step) == 11, sorted(step)      # cells<0..2> x {cs, ppm} + cells<3> + wind<0, 1> x {cs, ppm}
    for k, (reg, stack) in step.items():
        assert reg <= 64 and stack <= 128, (k, reg, stack)
    fz = [v for k, v in usage.items() if "k_fillz" in k]
    assert len(fz) == 1 and fz[0][0] <= 64 and fz[0][1] == 0, fz
