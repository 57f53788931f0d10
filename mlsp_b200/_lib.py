"""ctypes binding of libmlsp_b200.so (include/mlsp_b200.h).

The library is the product: if it is missing or a call fails this module raises -- there is no
CPU or PyTorch fallback for any op (BASELINE.json:north_star).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmlsp_b200.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_Z = ctypes.c_size_t
_F = ctypes.c_float

# name -> argtypes, in the order of include/mlsp_b200.h
SIGNATURES = {
    "mlsp_knn_f32": [_P, _I, _I, _I, _I, _P, _P, _Z, _I, _P],
    "mlsp_knn_tensor_debug": [_P, _I, _I, _I, _I, _P, _P, _Z, _P, _P],
    "mlsp_knn_tensor_timeline": [_P, _I, _I, _I, _I, _P, _P, _Z, _P, _I, _P],
    "mlsp_edge_gather_fwd": [_P, _P, _I, _I, _I, _I, _P, _P, _Z, _P],
    "mlsp_graph_feature_fwd": [_P, _I, _I, _I, _I, _P, _P, _P, _Z, _P],
    "mlsp_graph_feature_fwd_stage": [_P, _I, _I, _I, _I, _P, _P, _P, _Z, _I, _P],
    "mlsp_edge_gather_bwd": [_P, _P, _I, _I, _I, _I, _P, _P, _Z, _P],
    "mlsp_fps": [_P, _I, _I, _I, _P, _P, _P, _P],
    "mlsp_pcm_mix": [_P, _I, _I, _I, _P, _P, _P, _P, _P],
    "mlsp_region_assign_select": [_P, _L, _L, _L, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P],
    "mlsp_region_mask_scatter": [_P, _L, _L, _L, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "mlsp_ball_count": [_P, _L, _L, _L, _I, _I, _I, _F, _P, _P],
    "mlsp_ball_mask_scatter": [_P, _L, _L, _L, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    "mlsp_ball_count_labels": [_P, _I, _I, _F, _I, _I, _I, _I, _P, _P, _P],
    "mlsp_radius_search": [_P, _I, _I, _F, _I, _P, _P, _P],
    "mlsp_pca_normals": [_P, _P, _I, _I, _I, _P, _P, _P],
    "mlsp_target_structure": [_P, _I, _I, _I, _F, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "mlsp_chamfer_dir_fwd": [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _I, _I, _I, _P, _P, _P, _P, _Z, _P],
    "mlsp_chamfer_dir_bwd": [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _P, _I, _I, _P, _F, _P, _P, _P],
    "mlsp_reconstruction_loss_fwd": [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _I, _I, _P, _P, _P, _Z, _P],
    "mlsp_reconstruction_loss_bwd": [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _P, _I, _I, _P, _P, _P],
    "mlsp_edgeconv_reduce_fwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "mlsp_edgeconv_bn_coeffs": [_P, _P, _P, _I, ctypes.c_double, _F, _P, _P, _P, _P, _F, _P],
    "mlsp_edgeconv_apply_fwd": [_P, _P, _I, _I, _I, _F, _P, _P],
    "mlsp_edgeconv_bwd": [_P, _L, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P, _P, _Z, _P],
    "mlsp_edgeconv_weight_prep": [_P, _P, _P, _I, _I, _P, _P, _P, _P],
    "mlsp_edgeconv_weight_grad": [_P, _I, _P, _I, _I, _P, _P],
    "mlsp_max_mid_fwd": [_P, _L, _I, _I, _P, _P, _P],
    "mlsp_max_mid_bwd": [_P, _P, _L, _I, _I, _P, _P],
    "mlsp_max_row_fwd": [_P, _L, _I, _P, _P, _P],
    "mlsp_max_row_bwd": [_P, _P, _L, _I, _P, _P],
    "mlsp_scan_zbuffer": [_P, _I, _I, _P, _I, _P, _P, _P],
    "mlsp_gemm_f32": [_P, _I, _L, _L, _P, _I, _L, _L, _P, _I, _L, _L, _P, _I, _I, _I, _I, _P],
    "mlsp_bn_act_fwd": [_P, _P, _L, _I, _I, _I, _L, _L, _P, _P, _P, _P, _F, _F, _F, _P, _P, _P, _P],
    "mlsp_bn_act_bwd": [_P, _P, _P, _L, _I, _I, _I, _L, _L, _P, _P, _P, _P, _F, _P, _P, _P, _P],
    "mlsp_gemm_f32_timeline": [_P, _I, _L, _L, _P, _I, _L, _L, _P, _I, _L, _L, _P, _I, _I, _I, _I, _P, _P],
}

OP_KNN, OP_EDGE_FWD, OP_EDGE_BWD, OP_CHAMFER, OP_GRAPH_FEATURE, OP_EDGECONV_BWD = 1, 2, 3, 4, 5, 6
KNN_AUTO, KNN_EXACT_ONLY, KNN_TENSOR_ONLY, KNN_STATS = 0, 1, 2, 4


class MlspError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load the library once; raises MlspError if it has not been built (python -m mlsp_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MlspError(f"{LIB_PATH} not found: build it with `python -m mlsp_b200.build` "
                            "(there is no fallback implementation)")
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _I
        for hook in ("mlsp_fps_set_groups", "mlsp_fps_set_exclusive", "mlsp_knn_set_pdl"):     # void tuning hooks (include/mlsp_b200.h)
            getattr(L, hook).argtypes = [_I]
            getattr(L, hook).restype = None
        L.mlsp_version.restype = _I
        L.mlsp_last_error.restype = ctypes.c_char_p
        L.mlsp_workspace_bytes.argtypes = [_I, _I, _I, _I, _I]
        L.mlsp_workspace_bytes.restype = _Z
        _lib = L
    return _lib


calls = 0      # C-ABI entry points called so far (every one launches at least one kernel): bench.py's `gpu_launches` evidence


_fn: dict = {}


def call(name: str, *args) -> None:
    global calls
    fn = _fn.get(name)
    if fn is None:
        fn = _fn[name] = getattr(load(), name)
    calls += 1
    rc = fn(*args)
    if rc != 0:
        raise MlspError(f"{name} failed (code {rc}): {load().mlsp_last_error().decode()}")


def workspace_bytes(op: int, B: int, C: int, N: int, k: int) -> int:
    return int(load().mlsp_workspace_bytes(op, B, C, N, k))
