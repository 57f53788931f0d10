"""oracle/np_ops.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

numpy restatements of the integer / host-logic half of the MLSP target builder and of the
PCA normals.  float32 numpy arithmetic rounds once per written operation, so the op order
below *is* the spec.  Citations are reference file:line (VITA-Group/MLSP).
"""
from __future__ import annotations

import numpy as np

from . import ball_count, ball_row, density_count, knn

NREGIONS = 3          # utils/pc_utils.py:10
MIN_POINTS_BALL = 20  # utils/pc_utils.py:8
RADIUS = 0.5          # utils/pc_utils.py:9
MIN_PTS_VOXEL = 40    # MLSP/mlsp.py:27


def region_mean(num_regions=NREGIONS):
    """utils/pc_utils.py:13-30: centres of the n^3 voxels, region id = 9*qx + 3*qy + qz."""
    n = num_regions
    d = 2.0 / n
    c = np.array([1 - d * ((n - 1 - i) + 0.5) for i in range(n)])
    gx, gy, gz = np.meshgrid(c, c, c, indexing="ij")
    return np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)


def voxel_thresholds(n=NREGIONS):
    """The python-double bounds `-1 + q*d` of utils/pc_utils.py:57-62, cast to float32 exactly as
    torch does when a python scalar meets a float32 tensor in a comparison (probed: 0.1f == 0.1)."""
    d = 2 / n
    return np.array([np.float32(-1 + q * d) for q in range(n + 1)], dtype=np.float32)


def assign_region_to_point(X):
    """utils/pc_utils.py:33-73.  X (B,C>=3,N) float32 -> (B,N) int64.
    clamp(X, -0.99999999, 0.99999999): the constant is 1.0 in float32.  A point gets region
    9qx+3qy+qz only if all six STRICT comparisons hold; otherwise it keeps the initial 0."""
    X = np.asarray(X, dtype=np.float32)
    t = voxel_thresholds()
    Xc = np.clip(X[:, :3, :], np.float32(-1.0), np.float32(1.0))
    q = np.full(Xc.shape, -1, dtype=np.int64)
    for a in range(NREGIONS):
        inside = (t[a] < Xc) & (Xc < t[a + 1])
        q[inside] = a
    ok = (q >= 0).all(axis=1)
    rid = 9 * q[:, 0] + 3 * q[:, 1] + q[:, 2]
    return np.where(ok, rid, 0).astype(np.int64)


def region_hist(regions):
    """(B,N) -> (B,27) counts; the `torch.sum(regions[b] == i)` of MLSP/mlsp.py:39-41 for all i."""
    B = regions.shape[0]
    out = np.zeros((B, NREGIONS ** 3), np.int32)
    for b in range(B):
        out[b] = np.bincount(regions[b], minlength=NREGIONS ** 3)
    return out


def choose_region(counts, region_ids, min_pts=MIN_PTS_VOXEL):
    """MLSP/mlsp.py:37-50 with groups=1: first id in `region_ids` order with count >= min_pts; -1 if none."""
    B = counts.shape[0]
    chosen = np.full(B, -1, np.int64)
    for b in range(B):
        for i in region_ids:
            if counts[b, i] >= min_pts:
                chosen[b] = i
                break
    return chosen


def deform_input(X, lookup, DefRec_dist="volume_based_voxels", groups=1):
    """MLSP/mlsp.py:10-51 on numpy arrays, consuming the global numpy RNG in the reference's order.
    X (B,3,N) float32 is modified in place; returns (X, mask (B,3,N) float32)."""
    X = np.asarray(X)
    assert X.dtype == np.float32
    B, C, N = X.shape
    regions = assign_region_to_point(X)
    region_ids = np.random.permutation(NREGIONS ** 3)           # mlsp.py:28
    mask = np.zeros_like(X)
    if DefRec_dist == "volume_based_radius":
        for b in range(B):
            cnt = ball_count(X[b:b + 1], RADIUS ** 2, threads=1)[0]
            cand = np.nonzero(cnt >= MIN_POINTS_BALL)[0]         # pc_utils.py:96-99
            centre = np.random.choice(cand.squeeze())            # pc_utils.py:102
            flag = ball_row(X[b], centre, RADIUS ** 2)
            ind = np.nonzero(flag)[0]
            pts = np.random.multivariate_normal(X[b, :, centre], np.eye(3) * 0.001, len(ind)).T  # :122
            X[b][:, ind] = pts.astype(np.float32)
            mask[b][:3, ind] = 1
        return X, mask
    counts = region_hist(regions)
    lookup = np.asarray(lookup, dtype=np.float32)
    for b in range(B):
        iters = 0
        for i in region_ids:                                     # mlsp.py:37-50
            if counts[b, i] < MIN_PTS_VOXEL:
                continue
            iters += 1
            ind = regions[b] == i
            n = int(ind.sum())
            mask[b][:3, ind] = 1
            if DefRec_dist == "volume_based_voxels":
                pts = np.random.multivariate_normal(lookup[i], np.eye(3) * 0.001, n).T
                X[b][:3, ind] = pts.astype(np.float32)
            if iters >= groups:
                break
    return X, mask


def density_labels(cnt, num_cls, pergroup=2, shift=0):
    """MLSP/mlsp.py:254-266: row = clip(cnt - shift, 0, (num_cls-1)*pergroup);
    soft label = (onehot(floor(row/pg)) + onehot(ceil(row/pg))) / 2.  -> (float64 (B,N,num_cls), int64 (B,N))"""
    row = np.asarray(cnt, dtype=np.int64) - shift
    row = np.clip(row, 0, (num_cls - 1) * pergroup)
    lo = np.floor(row / pergroup).astype(np.int32)
    hi = np.ceil(row / pergroup).astype(np.int32)
    eye = np.identity(num_cls)
    return (eye[lo] + eye[hi]) / 2.0, row


def cal_density(batch_pts, radius, num_cls, pergroup=2, shift=0, K=100):
    """MLSP/mlsp.py:240-272 with the pcl radius search restated in mlsp_oracle.c:orc_density_count
    (PARITY UNPINNED: python-pcl is not runnable anywhere we can reach)."""
    cnt = density_count(batch_pts, radius, K)
    return density_labels(cnt, num_cls, pergroup, shift)


def pca_normals(xyz_bnc, near, return_gap=False):
    """kSearchNormalEstimation PointDA/trainer.py:173-188 (pcl NormalEstimation + KSearch), restated:
    k nearest neighbours incl. self (here: the reference's own knn formula) -> covariance of the
    neighbourhood about its mean -> unit eigenvector of the smallest eigenvalue, flipped towards the
    pcl default viewpoint (0,0,0): n.p <= 0.  fp64 throughout.  PARITY UNPINNED (python-pcl).
    xyz (B,N,3) float32 -> normals (B,N,3) float64 [, relative eigengap (B,N)]."""
    P = np.asarray(xyz_bnc, dtype=np.float32)
    B, N, _ = P.shape
    idx = knn(np.ascontiguousarray(P.transpose(0, 2, 1)), near)
    Pd = P.astype(np.float64)
    nb = Pd[np.arange(B)[:, None, None], idx]          # (B,N,near,3)
    mu = nb.mean(axis=2, keepdims=True)
    d = nb - mu
    cov = np.einsum("bnki,bnkj->bnij", d, d) / near
    w, v = np.linalg.eigh(cov)
    n = v[..., 0]
    flip = (n * Pd).sum(-1) > 0
    n = np.where(flip[..., None], -n, n)
    if return_gap:
        gap = (w[..., 1] - w[..., 0]) / np.maximum(w[..., 2], 1e-300)
        return n, gap
    return n


def rotation_matrix_3d():
    """rotate_point_cloud_3d's matrix (MLSP/mlsp.py:96-112): consumes np.random.rand(3)."""
    ang = np.random.rand(3) * 2 * np.pi
    c, s = np.cos(ang), np.sin(ang)
    r1 = np.array([[c[0], 0, s[0]], [0, 1, 0], [-s[0], 0, c[0]]])
    r2 = np.array([[1, 0, 0], [0, c[1], -s[1]], [0, s[1], c[1]]])
    r3 = np.array([[c[2], -s[2], 0], [s[2], c[2], 0], [0, 0, 1]])
    return np.matmul(np.matmul(r1, r2), r3)


def p_scan(pc, pixel_size, rot=None):
    """p_scan of MLSP/mlsp.py:66-94, vectorised: (N,3) float32 -> (scan_points, mask).  Bin = int((z'+1)/2*pixel*pixel +
    (y'+1)/2*pixel) of the rotated cloud on a (pixel+5)^2 grid (negative bins wrap like the Python list index they are);
    per bin the point with the largest x' survives, the first one on ties (the reference replaces on strict `>` only)."""
    pixel = int(2 / pixel_size)
    if rot is None:
        rot = rotation_matrix_3d()
    r = np.dot(pc.reshape((-1, 3)), rot)                                   # float32 . float64 -> float64, like the reference
    comp = ((r[:, 2] + 1) / 2 * pixel * pixel + (r[:, 1] + 1) / 2 * pixel).astype(np.int64)
    cells = (pixel + 5) * (pixel + 5)
    comp = np.where(comp < 0, comp + cells, comp)
    if ((comp < 0) | (comp >= cells)).any():
        raise IndexError("list index out of range")
    # first index of the per-bin maximum: sort by (bin, -x', index) and take the head of every bin
    order = np.lexsort((np.arange(len(comp)), -r[:, 0], comp))
    head = np.ones(len(comp), bool)
    head[1:] = comp[order][1:] != comp[order][:-1]
    keep = order[head]
    mask = np.ones_like(pc)
    mask[keep, :3] = 0.0
    out = np.zeros_like(pc)
    out[keep] = pc[keep]
    return out, mask


def scan_input(X, pixel_size=0.07):
    """scan_input of MLSP/mlsp.py:54-64 on a (B,N,3) float32 array: random.uniform for the pixel size, then per cloud p_scan."""
    import random
    pixel_size = random.uniform(0.045, 0.075)
    X = X.copy()
    mask = np.zeros_like(X)
    for b in range(X.shape[0]):
        X[b], mask[b] = p_scan(X[b], pixel_size)
    return X, mask
