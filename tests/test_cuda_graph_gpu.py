"""-m gpu: fv3_dyn_core(..., FV3_DYN_GRAPH) -- the whole call as one CUDA graph -- is bit-identical to the direct launches.

Covers what a replay must keep in step with the host: the fld <-> alt ping-pong pointers (an ODD number of substeps leaves them
swapped, so consecutive calls alternate between two captured graphs), the launch counter, the fall-back of the first call, a
change of bdt / n_split (new graph), the hydrostatic branch and the post-loop heating (flag-set B).
"""
import ctypes as C

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu
FIELDS = ("U", "V", "W", "DELZ", "PT", "DELP", "MFX", "MFY", "CX", "CY")


def _state(gc):
    return {t: {f: gc.eng[t].get(f) for f in FIELDS} for t in gc.tiles}


def _launches(gc):
    fn = gc.lib[0].fv3_launch_count
    fn.restype = C.c_longlong
    return [int(fn(gc.eng[t].ctx)) for t in gc.tiles]


@pytest.mark.parametrize("flagset,n_split,over", [("A", 2, None), ("A", 3, None), ("B", 2, None), ("A", 2, dict(hydrostatic=1))])
def test_graph_replay_is_bit_identical(flagset, n_split, over):
    case = H.Case(24, 6, flagset, state="baroclinic", flags_override=over)
    calls = [(400.0 * n_split, n_split)] * 4 + [(300.0 * n_split, n_split), (400.0 * n_split, n_split)]
    ref = H.CudaCube(case)
    gr = H.CudaCube(case)
    for n, (bdt, ns) in enumerate(calls):
        ref.dyn_core(bdt, ns)
        gr.dyn_core(bdt, ns, graph=True)    # call 0 falls back (first call), 1.. are captured / replayed
        a, b = _state(ref), _state(gr)
        for t in ref.tiles:
            for f in FIELDS:
                assert np.array_equal(a[t][f], b[t][f]), (n, t, f)
        assert _launches(ref) == _launches(gr), n
    ref.close(); gr.close()


def test_graph_flag_validation():
    case = H.Case(12, 3, "A", state="baroclinic")
    gc = H.CudaCube(case)
    fn = gc.lib[0].fv3_dyn_core
    fn.restype = C.c_int
    assert fn(gc.ctxs, len(gc.tiles), C.c_double(100.0), C.c_int(1), C.c_int(4)) == -2
    gc.close()
