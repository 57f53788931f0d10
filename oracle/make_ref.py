"""oracle/make_ref.py -- TEST INFRASTRUCTURE: stage the reference's OWN hot-path modules for the GPU box.

    python -m oracle.make_ref            (runs in the build container, where /root/reference is mounted)

The reference is pure Python, so "building" it means copying the modules that hold the hot path into
oracle/_ref/ -- git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so the unmodified files travel to the GPU box next to our own built artefacts -- and writing
stub packages for the third-party imports the hot path never reaches (pcl, timm,
knn_cuda, pointnet2_ops, termcolor, ...).  `bench.py --impl reference` and the `cpu_baseline` leg then time
THESE functions (oracle/ref_real.py) on the box's host cores: kind "reference".  Only the two python-pcl
calls (kd-tree cardinality, normal estimation) have no runnable reference anywhere; ref_real.py substitutes
the labelled dense-torch restatements of oracle/ref_torch.py for exactly those two.

Files staged (copied byte for byte; a manifest with their sha256 is written next to them):
    utils/pc_utils.py  MLSP/mlsp.py  MLSP/PCM.py  PointDA/model_utils.py  PointDA/Models.py  PointSegDA/Models.py
    utils/misc.py  utils/checkpoint.py  utils/log.py (imported by the above at module level)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
REF = os.environ.get("MLSP_REFERENCE", "/root/reference")

FILES = [
    "utils/pc_utils.py", "utils/misc.py", "utils/checkpoint.py", "utils/log.py",
    "MLSP/mlsp.py", "MLSP/PCM.py",
    "PointDA/model_utils.py", "PointDA/Models.py",
    "PointSegDA/Models.py",
]
PACKAGES = ["utils", "MLSP", "PointDA", "PointSegDA"]

# stubs for imports the hot path never executes (module level only)
STUBS = {
    "pcl/__init__.py": "# stub: python-pcl is not installable; only cal_density / kSearchNormalEstimation call it\n",
    "termcolor/__init__.py": "def colored(s, *a, **k):\n    return s\n",
    "timm/__init__.py": "",
    "timm/models/__init__.py": "",
    "timm/models/layers/__init__.py": ("import torch.nn as nn\n\n\nclass DropPath(nn.Identity):\n    def __init__(self, *a, **k):\n        super().__init__()\n\n\n"
                                       "def trunc_normal_(t, *a, **k):\n    return t\n"),
    "knn_cuda/__init__.py": "class KNN:\n    def __init__(self, *a, **k):\n        raise RuntimeError('knn_cuda is a stub (dead code on the hot path)')\n",
    "pointnet2_ops/__init__.py": "",
    "pointnet2_ops/pointnet2_utils.py": "def furthest_point_sample(*a, **k):\n    raise RuntimeError('pointnet2_ops is a stub')\n\n\ngather_operation = furthest_point_sample\n",
    "easydict/__init__.py": "class EasyDict(dict):\n    __getattr__ = dict.get\n",
    "h5py/__init__.py": "",
    "torchsummary/__init__.py": "def summary(*a, **k):\n    return None\n",
}


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "MLSP"))


def staged() -> bool:
    return os.path.exists(os.path.join(DST, "MANIFEST.json"))


def make(force: bool = False) -> str:
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF}")
    if staged() and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(DST, "src", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    for pkg in PACKAGES:
        init = os.path.join(DST, "src", pkg, "__init__.py")
        src_init = os.path.join(REF, pkg, "__init__.py")
        if os.path.exists(src_init):
            shutil.copyfile(src_init, init)
        elif not os.path.exists(init):
            open(init, "w").close()
    for rel, body in STUBS.items():
        p = os.path.join(DST, "stubs", rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(body)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"reference": REF, "files": manifest}, f, indent=1)
    return DST


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
