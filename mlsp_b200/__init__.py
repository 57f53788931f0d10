"""mlsp_b200 -- B200 (sm_100a) implementation of the MLSP data-parallel hot path.

Public surface = the reference's own function names (see ops.py) + `patch()` to rebind them
inside the reference modules.  Everything runs in hand-written CUDA kernels behind the C ABI
declared in include/mlsp_b200.h; there is no CPU fallback.
"""
from . import _lib, edgeconv, lazy, linear, ops, patch, pool, synth  # noqa: F401
from ._lib import MlspError  # noqa: F401
from .ops import (  # noqa: F401
    assign_region_to_point,
    ball_count,
    cal_density,
    calc_loss,
    chamfer_distance,
    collapse_to_point,
    deform_input,
    deform_input_begin,
    deform_input_finish,
    estimate_normals,
    farthest_point_sample,
    findindexs,
    findneareat_index,
    fps_from_start,
    get_graph_feature,
    knn,
    knn_tensor_debug,
    radius_search,
    reconstruction_loss,
    region_mean,
    scan_input,
    target_structure,
)

__version__ = "0.1.0"
