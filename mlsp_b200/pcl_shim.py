"""A minimal stand-in for the `pcl` module (python-pcl) backed by the B200 kernels.

The reference imports python-pcl at module scope (MLSP/mlsp.py:5, PointDA/trainer.py:18,
PointSegDA/trainer.py) and uses exactly this surface:

    cloud = pcl.PointCloud(); cloud.from_array(np.float32 (N,3))
    ne = cloud.make_NormalEstimation(); ne.set_SearchMethod(cloud.make_kdtree()); ne.set_KSearch(k)
    normals = ne.compute(); normals.size; normals[i][0]; normals.to_array() -> (N,4) [nx,ny,nz,curvature]
    kdtree = cloud.make_kdtree_flann(); kdtree.radius_search_for_cloud(cloud, r, K)   (only inside cal_density)

`install()` registers this module as `pcl` when the real one is absent, so the reference imports and its
per-cloud `kSearchNormalEstimation` loop (PointDA/trainer.py:173-188, :524-531) run on the GPU op
`mlsp_pca_normals`.  The batched fast path is `mlsp_b200.estimate_normals`.  `cal_density` is rebound as a whole
by `mlsp_b200.patch()` (one fused launch for the batch); an un-patched `cal_density` still works through
`radius_search_for_cloud` (GPU op `mlsp_radius_search`, one cloud per call like the reference's loop).
"""
from __future__ import annotations

import sys
import types

import numpy as np


class _Normals:
    def __init__(self, arr: np.ndarray):
        self._a = arr
        self.size = arr.shape[0]

    def __getitem__(self, i):
        return self._a[i]

    def to_array(self) -> np.ndarray:
        return self._a


class _KdTree:
    def __init__(self, cloud):
        self.cloud = cloud

    def radius_search_for_cloud(self, cloud, radius, K=100):
        """-> [ind (N,K) int32, sqdist (N,K) float32], rows zero-padded (python-pcl's return convention, consumed by
        MLSP/mlsp.py:250-253).  The tree's own cloud is searched around every point of `cloud`; the reference only
        ever passes a copy of the same points, which is what the GPU op supports."""
        import torch
        from . import ops
        q = cloud._pts if isinstance(cloud, PointCloud) else np.asarray(cloud, dtype=np.float32)
        if q.shape != self.cloud._pts.shape or not np.array_equal(q, self.cloud._pts):
            raise NotImplementedError("pcl shim: radius_search_for_cloud is provided for the reference's usage "
                                      "(query cloud == indexed cloud, MLSP/mlsp.py:243-250)")
        pts = torch.from_numpy(self.cloud._pts).cuda().unsqueeze(0)
        ind, sqd = ops.radius_search(pts, float(radius), int(K))
        return [ind[0].cpu().numpy(), sqd[0].cpu().numpy()]


class _NormalEstimation:
    def __init__(self, cloud):
        self.cloud = cloud
        self.k = None

    def set_SearchMethod(self, tree):
        pass

    def set_KSearch(self, k):
        self.k = int(k)

    def set_RadiusSearch(self, r):
        raise NotImplementedError("pcl shim: radius-search normals (radiusSearchNormalEstimation) are unused by the "
                                  "reference's training scripts and not provided")

    def compute(self) -> _Normals:
        import torch
        from . import ops
        if self.k is None:
            raise RuntimeError("pcl shim: set_KSearch(k) before compute()")
        pts = torch.from_numpy(self.cloud._pts).cuda().unsqueeze(0)
        n, curv = ops.estimate_normals(pts, self.k, return_curvature=True)
        out = torch.cat([n[0], curv[0].unsqueeze(1)], dim=1)
        return _Normals(out.cpu().numpy())


class PointCloud:
    def __init__(self, pts=None):
        self._pts = np.zeros((0, 3), np.float32)
        if pts is not None:
            self.from_array(np.asarray(pts, dtype=np.float32))

    def from_array(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        if arr.ndim != 2 or arr.shape[1] != 3:
            raise ValueError("pcl shim: expected an (N,3) float32 array")
        self._pts = arr

    def to_array(self):
        return self._pts

    @property
    def size(self):
        return self._pts.shape[0]

    def make_NormalEstimation(self):
        return _NormalEstimation(self)

    def make_kdtree(self):
        return _KdTree(self)

    def make_kdtree_flann(self):
        return _KdTree(self)


def install(force: bool = False) -> bool:
    """Register the shim as `pcl` unless a real python-pcl is importable.  Returns True if installed."""
    if not force:
        if "pcl" in sys.modules and not getattr(sys.modules["pcl"], "__mlsp_b200_shim__", False) \
                and hasattr(sys.modules["pcl"], "PointCloud"):
            return False
        try:
            import importlib.util
            if "pcl" not in sys.modules and importlib.util.find_spec("pcl") is not None:
                return False
        except (ImportError, ValueError):
            pass
    mod = types.ModuleType("pcl")
    mod.PointCloud = PointCloud
    mod.__mlsp_b200_shim__ = True
    sys.modules["pcl"] = mod
    return True
