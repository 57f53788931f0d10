"""oracle/ref_import.py -- TEST INFRASTRUCTURE: import shim for the *real* reference.

Only usable in the build container (where /root/reference is mounted); used by
oracle/gen_golden.py to produce tests/golden/ and by the container-only cross-checks in
tests/test_oracle_vs_reference.py (regenerates and byte-compares every fixture; skipped when the reference is absent, e.g. on the GPU box).
Third-party modules the reference imports but never reaches on the hot path are stubbed
(SURVEY.md section 8c).
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("MLSP_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "MLSP"))


def load():
    """Returns (pc_utils, mlsp, seg_models) -- the reference's own modules."""
    if not available():
        raise RuntimeError("reference checkout not present")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    sys.modules.setdefault("pcl", types.ModuleType("pcl"))      # MLSP/mlsp.py:5, only cal_density uses it
    import utils.pc_utils as pc_utils
    from MLSP import mlsp
    import PointSegDA.Models as seg_models                      # knn / get_graph_feature, no exotic deps
    return pc_utils, mlsp, seg_models
