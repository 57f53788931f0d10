import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


def knn_rank_check(x, idx, k, pd_ref=None):
    """Tie-aware check of a kNN result against fp64 ground truth of the reference formula.

    Returns (n_mismatch_ranks, n_uncertified).  A rank is a mismatch when the index differs from the
    fp64 stable ranking; it is *certified* when the fp64 pd values of the two candidates differ by less
    than 8 ulp(fp32) of the row's largest |term| -- i.e. a genuine rounding near-tie."""
    x = np.asarray(x, np.float64)
    B, C, N = x.shape
    bad = unc = 0
    for b in range(B):
        xx = (x[b] ** 2).sum(0)
        pd = 2 * x[b].T @ x[b] - xx[None, :] - xx[:, None]
        order = np.argsort(-pd, axis=1, kind="stable")[:, :k]
        diff = order != idx[b]
        if not diff.any():
            continue
        rows, ranks = np.nonzero(diff)
        scale = np.maximum(xx.max(), 1e-30)
        tol = 8 * np.finfo(np.float32).eps * scale
        for i, r in zip(rows, ranks):
            bad += 1
            if abs(pd[i, order[i, r]] - pd[i, idx[b, i, r]]) > tol:
                unc += 1
    return bad, unc
