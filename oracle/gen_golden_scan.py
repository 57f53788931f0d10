"""oracle/gen_golden_scan.py -- TEST INFRASTRUCTURE: tests/golden/scan_input.npz by the reference's OWN scan_input.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden_scan
MLSP/mlsp.py:54-94 (scan_input, p_scan, rotate_point_cloud_3d) is called unmodified; the only shim is `np.int = int`
(p_scan uses the alias numpy removed in 1.24 -- with this image's numpy 2.3 the reference function cannot run otherwise).
Seeds: random.seed / np.random.seed(SEED) right before the call; the clouds are mlsp_b200.synth surface clouds (inside the
unit ball, as the reference's loaders normalise them), one of them with duplicated points (ties inside a bin)."""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mlsp_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 23


def main():
    _, mlsp, _ = ref_import.load()
    if not hasattr(np, "int"):
        np.int = int                                            # noqa: NPY001  (see the module docstring)
    X = synth.surface_clouds(4, 1024, 91).permute(0, 2, 1).contiguous()          # (B,N,3)
    X[3, 512:] = X[3, :512]                                     # duplicates: equal x' inside a bin, the first index must win
    random.seed(SEED)
    np.random.seed(SEED)
    out, mask = mlsp.scan_input(X.clone(), torch.device("cpu"))
    np.savez_compressed(os.path.join(OUT, "scan_input.npz"), X=X.numpy(), seed=SEED, out=out.numpy(), mask=mask.numpy())
    print("scan_input.npz", os.path.getsize(os.path.join(OUT, "scan_input.npz")), "kept per cloud", (mask[:, :, 0] == 0).sum(1).tolist())


if __name__ == "__main__":
    main()
