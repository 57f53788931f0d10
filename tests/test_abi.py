"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/mlsp_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mlsp_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mlsp_b200.h")).read()
    return sorted(set(re.findall(r"MLSP_API[^;(]*?(mlsp_\w+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("mlsp_knn_f32", "mlsp_edge_gather_fwd", "mlsp_edge_gather_bwd", "mlsp_fps",
                 "mlsp_region_assign_select", "mlsp_region_mask_scatter", "mlsp_ball_count",
                 "mlsp_ball_mask_scatter", "mlsp_ball_count_labels", "mlsp_pca_normals",
                 "mlsp_chamfer_dir_fwd", "mlsp_chamfer_dir_bwd", "mlsp_workspace_bytes", "mlsp_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_ctypes_signatures_cover_header(lib):
    from mlsp_b200 import _lib
    compute = [s for s in declared_symbols() if s not in ("mlsp_version", "mlsp_last_error", "mlsp_workspace_bytes",
                                                          "mlsp_fps_set_groups", "mlsp_fps_set_exclusive",        # void tuning hooks
                                                          "mlsp_bn_scratch_bytes", "mlsp_knn_set_pdl")]
    assert sorted(compute) == sorted(_lib.SIGNATURES)


def test_version_and_workspace(lib):
    from mlsp_b200 import _lib
    assert lib.mlsp_version() >= 100
    assert _lib.workspace_bytes(_lib.OP_KNN, 32, 3, 1024, 20) >= 32 * 1024 * 4
    assert _lib.workspace_bytes(_lib.OP_EDGE_FWD, 32, 64, 1024, 20) >= 32 * 64 * 1024 * 4
    assert _lib.workspace_bytes(99, 1, 1, 1, 1) == 0


def test_argument_errors_without_gpu(lib):
    """Shape validation happens before any CUDA call, so error codes are testable on CPU."""
    rc = lib.mlsp_knn_f32(None, 1, 3, 16, 4, None, None, 0, 0, None)
    assert rc == 1 and b"null" in lib.mlsp_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.mlsp_knn_f32(p, 1, 3, 16, 40, p, p, 1 << 20, 0, None)
    assert rc == 1 and b"out of range" in lib.mlsp_last_error()
    rc = lib.mlsp_knn_f32(p, 70000, 3, 1024, 20, p, p, 1 << 20, 0, None)          # B > 65535 (grid y): MLSP_EUNSUPPORTED
    assert rc == 2
    rc = lib.mlsp_knn_f32(p, 1, 3, 1024, 100, p, p, 16, 0, None)                  # k > 64 is accepted: the next check is the workspace
    assert rc == 4 and b"workspace" in lib.mlsp_last_error()


def test_edgeconv_argument_errors_without_gpu(lib):
    """The EdgeConv entry points validate shapes / pointers / workspace before any CUDA call (codes of mlsp_b200.h)."""
    buf = ctypes.create_string_buffer(256)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.mlsp_edgeconv_reduce_fwd(None, p, 1, 16, 8, 4, p, p, None, None, None) == 1 and b"null" in lib.mlsp_last_error()
    assert lib.mlsp_edgeconv_reduce_fwd(p, p, 1, 16, 6, 4, p, p, None, None, None) == 2          # O % 4 != 0
    assert lib.mlsp_edgeconv_reduce_fwd(p, p, 1, 1024, 8, 300, p, p, None, None, None) == 2      # k > 255: slots are bytes
    assert lib.mlsp_edgeconv_reduce_fwd(p, p, 1, 16, 8, 40, p, p, None, None, None) == 1         # k > N
    assert lib.mlsp_edgeconv_reduce_fwd(p, p, 1, 16, 8, 4, p, p, p, None, None) == 1             # rowsum without stats
    assert lib.mlsp_edgeconv_bn_coeffs(p, None, None, 8, 64.0, 1e-5, p, p, None, None, 0.1, None) == 1   # running_mean without running_var
    assert lib.mlsp_edgeconv_apply_fwd(p, p, 0, 16, 8, 0.2, p, None) == 1
    args = (p, 16 * 8, p, p, p, p, None, p, 1, 16, 8, 4, 0.2)
    assert lib.mlsp_edgeconv_bwd(*args, 1, p, p, p, 1 << 20, None) == 1 and b"rowsum" in lib.mlsp_last_error()   # bn_train needs rowsum
    assert lib.mlsp_edgeconv_bwd(*args, 0, p, None, p, 16, None) == 4                             # workspace too small
    assert lib.mlsp_edgeconv_bwd(p, 8, p, p, p, p, None, p, 1, 16, 8, 4, 0.2, 0, p, None, p, 1 << 20, None) == 1   # g batch stride < O*N
    from mlsp_b200 import _lib
    assert _lib.workspace_bytes(_lib.OP_EDGECONV_BWD, 32, 64, 1024, 20) >= 3 * 32 * 1024 * 64 * 4


def test_ops_refuse_cpu_tensors():
    import torch
    import mlsp_b200 as M
    with pytest.raises(M.MlspError):
        M.knn(torch.zeros(1, 3, 8), 2)
    with pytest.raises(M.MlspError):
        M.get_graph_feature(torch.zeros(1, 3, 8), None, k=2)
    with pytest.raises(M.MlspError):
        M.reconstruction_loss(torch.zeros(1, 8, 3), torch.zeros(1, 3, 8), torch.zeros(1, 3, 8))


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (it is the checker)."""
    pkg = os.path.join(ROOT, "mlsp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
