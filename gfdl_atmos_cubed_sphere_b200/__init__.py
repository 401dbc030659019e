"""B200-native FV3 acoustic-dynamics hot path (c_sw / d_sw / fv_tp_2d / Riemann solvers behind the C ABI of include/fv3_dyncore.h).

    from gfdl_atmos_cubed_sphere_b200 import Case, CudaCube
    cube = CudaCube(Case(96, 79, "A", state="baroclinic"))     # six faces on cuda:0, linked for the halo exchange
    cube.dyn_core(bdt=225.0, n_split=8)

The CUDA library (csrc/libfv3_b200.so) is loaded lazily by `abi.load_library()`; there is no CPU fallback.
"""
from . import abi  # noqa: F401
from .cube import Case, CudaCube, FLAGSETS  # noqa: F401
from .parallel import tiles_of_rank, tile_rank_map  # noqa: F401
