"""-m gpu: the vertical remap with ppm_profile (abs(kord) <= 7, fv_operators.F90:1382-1723) against the oracle through the C ABI.

The PPM instantiations of the remap kernels (csrc/remap.cu: k_remap_cells<., true>, k_remap_wind<., true>, k_remap_work_q<true>) were
written late in round 2, pinned without a device first --
  * their column code (csrc/remap_col.cuh: ppm_profile / ppm_limiters, __host__ __device__) runs on the host and equals the oracle
    BIT FOR BIT for every scheme 3..7 in every mode (tests/test_host_remap.py);
  * the oracle side is held to an oracle-independent known answer (tests/test_remap_oracle.py: a quadratic profile is remapped to
    rounding error in the interior);
  * the instruction streams of the kernels of the schemes 8..15 are unchanged (the PPM code lives in separate instantiations)
-- and then run on a B200 with the last seconds of the round's GPU budget: 9 passed (profiles/r2/r2g_ppm_remap_pytest.txt), so
they are plain tests (no xfail guard)."""
import pytest

import harness as H
from test_remap_gpu import TOL, _assert, _pair, _regions

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kord", [3, 4, 5, 6, 7])
def test_ppm_column_operators_match_the_oracle(kord):
    case, oc, gc = _pair()
    b = case.bounds
    reg = {"WORK_Q": (b["is_"], b["ie"], b["js"], b["je"])}
    for mode, iv in [(0, 1), (1, 1), (1, -1), (1, -2), (2, 0), (0, 0), (1, 2)]:
        for t in (2, 5):
            eo, eg = oc.eng[t], gc.eng[t]
            q0 = eo.get("WORK_Q")
            for e in (eo, eg):
                e.call("remap_work_q", mode, iv, kord, 1.0 if mode == 0 else 0.0)
            _assert(H.compare(eo, eg, reg), TOL)
            for e in (eo, eg):
                e.put("WORK_Q", q0)
    oc.close(); gc.close()


@pytest.mark.parametrize("kord_tm,kord,last,tracer,hydro", [(-7, 7, 0, 1, 0), (4, 4, 1, 1, 0), (-6, 6, 1, 0, 1), (-9, 5, 0, 1, 0)])
def test_ppm_lagrangian_to_eulerian_matches_the_oracle(kord_tm, kord, last, tracer, hydro):
    case, oc, gc = _pair(hydrostatic=hydro)
    reg = _regions(case.bounds, bool(hydro))
    for t in oc.tiles:
        for e in (oc.eng[t], gc.eng[t]):
            e.call("lagrangian_to_eulerian", last, kord, kord, kord_tm, tracer, kord)
        _assert(H.compare(oc.eng[t], gc.eng[t], reg), TOL)
    oc.close(); gc.close()
