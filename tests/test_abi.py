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
    compute = [s for s in declared_symbols() if s not in ("mlsp_version", "mlsp_last_error", "mlsp_workspace_bytes")]
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
    rc = lib.mlsp_knn_f32(p, 1, 3, 1024, 100, p, p, 1 << 20, 0, None)
    assert rc == 2


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
