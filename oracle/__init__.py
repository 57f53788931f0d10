"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the MLSP hot path used as the parity checker for the sm_100a kernels.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; nothing under ``mlsp_b200/`` does.

Three layers:

* ``mlsp_oracle.c`` (ctypes, this module)  -- exact arithmetic spec for the FP-compare ->
  integer ops (kNN, FPS, ball / cardinality counts, Chamfer argmin) and their FP outputs.
* ``np_ops.py``  -- numpy restatements of the byte/integer host logic (voxel region ids,
  histogram + region choice, soft cardinality labels) and of the PCA normals (fp64 eigh).
* ``ref_torch.py`` -- a pure-torch port that keeps the reference's *op composition*
  (matmul + topk, repeat + norm + min, the FPS python loop ...); it is what
  ``bench.py --impl reference`` times on the host cores, because the Python reference at
  /root/reference cannot travel to the GPU box.

Pinning: ``oracle/gen_golden.py`` imports the reference's own functions in the build
container and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every layer
against them.  The two python-pcl backed pieces (cardinality a6, normals a7) have no
runnable reference anywhere (python-pcl is un-vendored, unpinned and not installable):
for those two the oracle is **parity unpinned**.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build(force: bool = False) -> None:
    """Compile mlsp_oracle.c with the committed Makefile (gcc only, no reference sources)."""
    so = os.path.join(_BUILD, "libmlsp_oracle.so")
    src = os.path.join(_HERE, "mlsp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])


def _cpu_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in (line + " ")
    except OSError:
        pass
    return False


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        name = "libmlsp_oracle.so" if _cpu_has_fma() else "libmlsp_oracle_nofma.so"
        L = ctypes.CDLL(os.path.join(_BUILD, name))
        c_f = ctypes.POINTER(ctypes.c_float)
        c_d = ctypes.POINTER(ctypes.c_double)
        c_l = ctypes.POINTER(ctypes.c_int64)
        c_i = ctypes.POINTER(ctypes.c_int32)
        c_b = ctypes.POINTER(ctypes.c_uint8)
        I = ctypes.c_int
        L.orc_knn.argtypes = [c_f, I, I, I, I, c_l, c_f]
        L.orc_knn_row_f64.argtypes = [c_f, I, I, I, c_d]
        L.orc_edge_gather.argtypes = [c_f, c_l, I, I, I, I, c_f]
        L.orc_edge_gather_bwd.argtypes = [c_f, c_l, I, I, I, I, c_d]
        L.orc_fps.argtypes = [c_f, I, I, I, c_l, c_l, c_f]
        L.orc_ball_count.argtypes = [c_f, I, I, ctypes.c_float, c_i]
        L.orc_ball_row.argtypes = [c_f, I, ctypes.c_float, I, c_b]
        L.orc_density_count.argtypes = [c_f, I, I, ctypes.c_float, I, c_i]
        L.orc_radius_search.argtypes = [c_f, I, I, ctypes.c_float, I, c_i, c_f]
        L.orc_chamfer_dir.argtypes = [c_f, c_f, c_f, I, I, c_f, c_l]
        L.orc_chamfer_dir.restype = ctypes.c_double
        L.orc_reconstruction_loss.argtypes = [c_f, c_f, c_f, I, I, c_d]
        L.orc_reconstruction_loss.restype = ctypes.c_double
        for fn in ("orc_knn", "orc_knn_row_f64", "orc_edge_gather", "orc_edge_gather_bwd", "orc_fps",
                   "orc_ball_count", "orc_ball_row", "orc_density_count", "orc_radius_search"):
            getattr(L, fn).restype = I
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _d(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _l(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _threads(threads):
    return max(1, int(threads if threads else (os.cpu_count() or 1)))


def _over_batches(B, threads, fn):
    """Run fn(b0, b1) over contiguous batch slices on a thread pool (ctypes drops the GIL)."""
    t = min(_threads(threads), B)
    if t <= 1:
        fn(0, B)
        return
    cuts = np.linspace(0, B, t + 1).astype(int)
    with ThreadPoolExecutor(t) as ex:
        list(ex.map(lambda ab: fn(int(ab[0]), int(ab[1])), zip(cuts[:-1], cuts[1:])))


def _check(rc, what):
    if rc != 0:
        raise ValueError(f"oracle {what}: bad arguments (rc={rc})")


# --------------------------------------------------------------------------- a1
def knn(x, k, return_pd=False, threads=None):
    """(B,C,N) float32 -> (B,N,k) int64; ranking pd desc, ties lowest index."""
    x = _f32(x)
    B, C, N = x.shape
    idx = np.empty((B, N, k), np.int64)
    pd = np.empty((B, N, k), np.float32) if return_pd else None

    def run(b0, b1):
        _check(lib().orc_knn(_f(x[b0:b1]), b1 - b0, C, N, k, _l(idx[b0:b1]),
                             _f(pd[b0:b1]) if return_pd else None), "knn")

    _over_batches(B, threads, run)
    return (idx, pd) if return_pd else idx


def knn_row_f64(xb, i):
    """fp64 pd row of one point (near-tie certificate)."""
    xb = _f32(xb)
    C, N = xb.shape
    row = np.empty(N, np.float64)
    lib().orc_knn_row_f64(_f(xb), C, N, int(i), _d(row))
    return row


# --------------------------------------------------------------------------- a2
def edge_gather(x, idx, threads=None):
    """x (B,C,N), idx (B,N,k) -> channels-last storage (B,N,k,2C) float32."""
    x = _f32(x)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    B, C, N = x.shape
    k = idx.shape[2]
    out = np.empty((B, N, k, 2 * C), np.float32)

    def run(b0, b1):
        _check(lib().orc_edge_gather(_f(x[b0:b1]), _l(idx[b0:b1]), b1 - b0, C, N, k, _f(out[b0:b1])),
               "edge_gather")

    _over_batches(B, threads, run)
    return out


def edge_gather_bwd(g, idx, C, threads=None):
    """g (B,N,k,2C) channels-last grad -> grad_x (B,C,N) float64."""
    g = _f32(g)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    B, N, k, _ = g.shape
    gx = np.empty((B, C, N), np.float64)

    def run(b0, b1):
        _check(lib().orc_edge_gather_bwd(_f(g[b0:b1]), _l(idx[b0:b1]), b1 - b0, C, N, k, _d(gx[b0:b1])),
               "edge_gather_bwd")

    _over_batches(B, threads, run)
    return gx


# --------------------------------------------------------------------------- a3
def fps(xyz, npoint, start, threads=None):
    """xyz (B,3,N), start (B,) -> centroids (B,npoint) int64, vals (B,3,npoint) float32."""
    xyz = _f32(xyz)
    start = np.ascontiguousarray(start, dtype=np.int64)
    B, C, N = xyz.shape
    assert C == 3
    cen = np.empty((B, npoint), np.int64)
    vals = np.empty((B, 3, npoint), np.float32)

    def run(b0, b1):
        _check(lib().orc_fps(_f(xyz[b0:b1]), b1 - b0, N, npoint, _l(start[b0:b1]), _l(cen[b0:b1]),
                             _f(vals[b0:b1])), "fps")

    _over_batches(B, threads, run)
    return cen, vals


# --------------------------------------------------------------------------- a5
def ball_count(x, r2=0.25, threads=None):
    """x (B,3,N) -> (B,N) int32 in-ball counts with the collapse_to_point formula."""
    x = _f32(x)
    B, C, N = x.shape
    assert C == 3
    cnt = np.empty((B, N), np.int32)

    def run(b0, b1):
        _check(lib().orc_ball_count(_f(x[b0:b1]), b1 - b0, N, np.float32(r2), _i(cnt[b0:b1])), "ball_count")

    _over_batches(B, threads, run)
    return cnt


def ball_row(xb, centre, r2=0.25):
    xb = _f32(xb)
    N = xb.shape[1]
    flag = np.empty(N, np.uint8)
    lib().orc_ball_row(_f(xb), N, np.float32(r2), int(centre), flag.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return flag.astype(bool)


# --------------------------------------------------------------------------- a6
def density_count(pts, radius, K=100, threads=None):
    """pts (B,N,3) -> (B,N) int32 raw cardinality (before shift / clip)."""
    pts = _f32(pts)
    B, N, _ = pts.shape
    r2 = np.float32(float(radius) * float(radius))  # pcl: static_cast<float>(radius*radius) in double
    cnt = np.empty((B, N), np.int32)

    def run(b0, b1):
        _check(lib().orc_density_count(_f(pts[b0:b1]), b1 - b0, N, r2, int(K), _i(cnt[b0:b1])), "density_count")

    _over_batches(B, threads, run)
    return cnt


def radius_search(pts, radius, K=100, threads=None):
    """pts (B,N,3) -> (ind (B,N,K) int32, sqdist (B,N,K) float32), rows zero-padded: the restated
    pcl radius_search_for_cloud whose `(ind != 0).sum(-1)` is density_count."""
    pts = _f32(pts)
    B, N, _ = pts.shape
    r2 = np.float32(float(radius) * float(radius))
    ind = np.empty((B, N, int(K)), np.int32)
    sqd = np.empty((B, N, int(K)), np.float32)

    def run(b0, b1):
        _check(lib().orc_radius_search(_f(pts[b0:b1]), b1 - b0, N, r2, int(K), _i(ind[b0:b1]), _f(sqd[b0:b1])),
               "radius_search")

    _over_batches(B, threads, run)
    return ind, sqd


# --------------------------------------------------------------------------- a9 / a10
def chamfer_dir(p1, p2, mask, threads=None):
    """One direction of the masked Chamfer: returns (sum_b S_b, rowmin (B,N), argmin (B,N))."""
    p1, p2, mask = _f32(p1), _f32(p2), _f32(mask)
    B, N, _ = p1.shape
    rm = np.empty((B, N), np.float32)
    am = np.empty((B, N), np.int64)
    parts = []

    def run(b0, b1):
        parts.append(lib().orc_chamfer_dir(_f(p1[b0:b1]), _f(p2[b0:b1]), _f(mask[b0:b1]), b1 - b0, N,
                                           _f(rm[b0:b1]), _l(am[b0:b1])))

    _over_batches(B, threads, run)
    return float(np.sum(parts)), rm, am


def reconstruction_loss(pred, gold_bcn, mask_bcn, with_grad=True):
    """pred (B,N,3); gold, mask (B,3,N) as the reference passes them. -> (loss, grad_pred float64)."""
    pred = _f32(pred)
    gold = _f32(np.transpose(np.asarray(gold_bcn), (0, 2, 1)))
    mask = _f32(np.asarray(mask_bcn)[:, 0, :])
    B, N, _ = pred.shape
    grad = np.empty((B, N, 3), np.float64) if with_grad else None
    loss = lib().orc_reconstruction_loss(_f(pred), _f(gold), _f(mask), B, N, _d(grad) if with_grad else None)
    return float(loss), grad
