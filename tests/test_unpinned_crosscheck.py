"""Independent cross-checks for the two rows whose reference lives in python-pcl (a6 cardinality, a7 normals) and
cannot be run anywhere we can reach -- "parity unpinned" (SURVEY.md 8c).  The oracle restatements (oracle/np_ops.py,
oracle/mlsp_oracle.c) are compared here with implementations that share NO code with them:

* a6: scipy's cKDTree (a k-d tree radius query in float64, like pcl's FLANN kd-tree) for the in-ball sets, then the
  reference's own label arithmetic (MLSP/mlsp.py:252-266) on top;
* a7: the algorithm pcl::NormalEstimation documents -- single-pass float32 mean / covariance of the k nearest neighbours
  (computeMeanAndCovarianceMatrix), smallest-eigenvalue eigenvector, flipped towards the viewpoint (0,0,0) -- with the
  neighbourhoods from cKDTree.query.

Results are MISMATCH RATES, asserted against small bounds: the remaining differences are float32-vs-float64 boundary
decisions (|d - r| below rounding) and near-degenerate covariances, listed in the assertion messages.
The GPU twins of these checks are in tests/test_gpu_parity.py (test_cal_density_vs_ckdtree, test_normals_vs_pcl_style)."""
import numpy as np
import pytest

scipy_spatial = pytest.importorskip("scipy.spatial")

from mlsp_b200 import synth          # noqa: E402
from oracle import np_ops            # noqa: E402


def ckdtree_density_rows(pts, radius, num_cls, pergroup, shift, K):
    """cal_density's `row` (MLSP/mlsp.py:240-258) from cKDTree: in-ball = d < r (FLANN's strict test; cKDTree's ball is closed,
    so exact-boundary points are removed explicitly), at most the K nearest, neighbour index 0 dropped (`ind != 0` on the
    zero-padded index array, mlsp.py:252)."""
    B, N, _ = pts.shape
    rows = np.zeros((B, N), np.int64)
    for b in range(B):
        P = pts[b].astype(np.float64)
        tree = scipy_spatial.cKDTree(P)
        d, j = tree.query(P, k=min(K, N), distance_upper_bound=radius)       # sorted by distance, inf-padded
        inside = np.isfinite(d) & (d < radius)
        cnt = inside.sum(1) - (inside & (j == 0)).sum(1)
        rows[b] = np.clip(cnt - shift, 0, (num_cls - 1) * pergroup)
    return rows


def pcl_style_normals(pts, near):
    """pcl::NormalEstimation with KSearch(near), as documented: float32 single-pass centroid + covariance, eigen-solve,
    flip towards the origin.  -> normals (B,N,3) float32, relative eigengap (B,N)."""
    B, N, _ = pts.shape
    out = np.zeros((B, N, 3), np.float32)
    gap = np.zeros((B, N))
    for b in range(B):
        P = pts[b].astype(np.float32)
        _, idx = scipy_spatial.cKDTree(P.astype(np.float64)).query(P.astype(np.float64), k=near)
        nb = P[idx]                                                          # (N,near,3) float32
        s1 = nb.sum(1, dtype=np.float32) / np.float32(near)                   # accu[6..8] / n
        s2 = np.einsum("nki,nkj->nij", nb, nb, dtype=np.float32) / np.float32(near)
        cov = (s2 - s1[:, :, None] * s1[:, None, :]).astype(np.float32)       # E[xx^T] - mean mean^T in float32
        w, v = np.linalg.eigh(cov.astype(np.float64))
        n = v[:, :, 0]
        n = np.where(((n * P).sum(-1) > 0)[:, None], -n, n)
        out[b] = n.astype(np.float32)
        gap[b] = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2], 1e-300)
    return out, gap


@pytest.mark.parametrize("B,N,radius,num_cls,pergroup,shift,K", [(4, 1024, 0.13, 16, 2, 0, 100), (2, 2048, 0.091, 16, 5, 10, 100),
                                                               (2, 1024, 0.4, 16, 2, 0, 100)])
def test_cardinality_oracle_vs_ckdtree(B, N, radius, num_cls, pergroup, shift, K):
    pts = synth.surface_clouds(B, N, 21).permute(0, 2, 1).contiguous().numpy()
    _, rows = np_ops.cal_density(pts, radius, num_cls, pergroup, shift, K)
    ref = ckdtree_density_rows(pts, radius, num_cls, pergroup, shift, K)
    mism = float((rows != ref).mean())
    worst = int(np.abs(rows - ref).max())
    # float32 squared distances against float64 distances: only points within rounding of the sphere can differ, by one count
    assert mism <= 2e-3 and worst <= 1, (mism, worst)


@pytest.mark.parametrize("B,N,near", [(4, 1024, 20), (2, 2048, 10)])
def test_normals_oracle_vs_pcl_style(B, N, near):
    pts = synth.surface_clouds(B, N, 22).permute(0, 2, 1).contiguous().numpy()
    ours, gap = np_ops.pca_normals(pts, near, return_gap=True)
    ref, gap32 = pcl_style_normals(pts, near)
    cos = np.abs((ours * ref.astype(np.float64)).sum(-1))
    well = (gap > 1e-2) & (gap32 > 1e-2)
    # float32 single-pass covariance loses ~|mean|^2 * 2^-24 against the spread of a 20-point patch: 1e-3 is what it supports
    frac_1e3 = float((1.0 - cos[well] <= 1e-3).mean())
    frac_1e5 = float((1.0 - cos[well] <= 1e-5).mean())
    assert well.mean() > 0.9 and frac_1e3 >= 0.995, (float(well.mean()), frac_1e3, frac_1e5)
    assert ((ours * pts).sum(-1) <= 1e-12).all() and ((ref * pts).sum(-1) <= 1e-6).all()   # both face the origin
