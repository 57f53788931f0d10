"""Drop-in boundary: rebind the reference's hot-path names to the B200 ops.

The reference has no plugin or FFI layer; its boundary is Python name binding.  `patch()` walks the
modules that define or captured the hot-path functions and replaces the attributes in place, so
Models.py / mlsp.py / trainer.py keep calling `get_graph_feature(...)`, `mlsp.deform_input(...)`,
`mlsp.calc_loss(...)`, `pc_utils.farthest_point_sample(...)` unchanged.

    import mlsp_b200
    mlsp_b200.patch.install_pcl_shim()        # before the reference imports `pcl`
    import PointDA.Models, MLSP.mlsp, utils.pc_utils
    mlsp_b200.patch.patch()                   # rebinds every captured name, returns what it touched

`patch(fuse_edgeconv=True)` binds `get_graph_feature` to mlsp_b200.lazy.get_graph_feature instead: the graph feature is
deferred, and a following `conv_2d(...)` / Conv2d stack + `.max(dim=-1)` runs as one EdgeConv layer without the
(B,2C,N,k) tensor -- still with the reference's model code unchanged.
"""
from __future__ import annotations

import sys

from . import lazy, ops, pcl_shim

# module name -> {attribute: replacement}.  Both `PointDA.model_utils` and the top-level `model_utils`
# exist as distinct module objects in the reference (PointDA/Models.py:4 vs :10), and Models.py binds
# get_graph_feature by `from ... import`, so each namespace is listed.
_NEIGHBOURHOOD = {"knn": ops.knn, "get_graph_feature": ops.get_graph_feature}
TARGETS = {
    "PointDA.model_utils": _NEIGHBOURHOOD,
    "model_utils": _NEIGHBOURHOOD,
    "PointDA.Models": _NEIGHBOURHOOD,
    "Models": _NEIGHBOURHOOD,
    "PointSegDA.Models": _NEIGHBOURHOOD,
    "utils.pc_utils": {
        "farthest_point_sample": ops.farthest_point_sample,
        "assign_region_to_point": ops.assign_region_to_point,
        "collapse_to_point": ops.collapse_to_point,
    },
    "MLSP.mlsp": {
        "deform_input": ops.deform_input,
        "scan_input": ops.scan_input,
        "chamfer_distance": ops.chamfer_distance,
        "reconstruction_loss": ops.reconstruction_loss,
        "findneareat_index": ops.findneareat_index,
        "findindexs": ops.findindexs,
        "cal_density": ops.cal_density,
    },
}


def install_pcl_shim(force: bool = False) -> bool:
    return pcl_shim.install(force)


def patch(modules=None, strict: bool = False, fuse_edgeconv: bool = False):
    """Rebind the hot-path names in every already-imported reference module (or in `modules`, a dict
    name -> module object).  Only attributes the module already has are replaced.  Returns a list of
    "module.attr" strings; with strict=True raises if nothing was patched.  fuse_edgeconv=True: get_graph_feature
    returns the deferred feature of mlsp_b200.lazy (EdgeConv layers run without the edge tensor)."""
    touched = []
    originals = {}
    for modname, table in TARGETS.items():
        if fuse_edgeconv and "get_graph_feature" in table:
            table = dict(table, get_graph_feature=lazy.get_graph_feature)
        mod = (modules or {}).get(modname) or sys.modules.get(modname)
        if mod is None:
            continue
        for attr, fn in table.items():
            if hasattr(mod, attr) and getattr(mod, attr) is not fn:
                originals[(modname, attr)] = getattr(mod, attr)
                setattr(mod, attr, fn)
                touched.append(f"{modname}.{attr}")
    for key, fn in originals.items():                 # a second patch() (e.g. fuse_edgeconv=True after a plain one) must not
        patch.originals.setdefault(key, fn)           # forget the reference's own function
    if strict and not touched:
        raise RuntimeError("mlsp_b200.patch: no reference module is imported yet")
    return touched


patch.originals = {}


def unpatch():
    """Restore whatever patch() replaced."""
    for (modname, attr), fn in list(patch.originals.items()):
        mod = sys.modules.get(modname)
        if mod is not None:
            setattr(mod, attr, fn)
    patch.originals.clear()
