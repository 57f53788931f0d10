"""Container-only cross-check: every committed fixture regenerates byte for byte from the committed recipes
(oracle/gen_golden.py, oracle/gen_golden_edgeconv.py, oracle/gen_golden_activations.py), which call the REFERENCE'S OWN functions.  Skipped where the
reference checkout is absent (the GPU box); arrays are compared, not the zip containers (timestamps)."""
import importlib
import os

import numpy as np
import pytest

from oracle import ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")


@pytest.mark.parametrize("modname", ["oracle.gen_golden", "oracle.gen_golden_edgeconv", "oracle.gen_golden_activations", "oracle.gen_golden_scan"])
def test_fixtures_regenerate_identically(tmp_path, modname, monkeypatch):
    mod = importlib.import_module(modname)
    monkeypatch.setattr(mod, "OUT", str(tmp_path))
    mod.main()
    made = sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz"))
    assert made, "the generator wrote nothing"
    for f in made:
        new, old = np.load(os.path.join(tmp_path, f)), np.load(os.path.join(GOLDEN, f))
        assert sorted(new.files) == sorted(old.files), f
        for key in new.files:
            same = np.array_equal(new[key], old[key], equal_nan=new[key].dtype.kind == "f")   # chamfer_empty pins a NaN loss
            assert same, (f, key)


def test_staged_reference_is_the_reference():
    """oracle/_ref (what `bench.py --impl reference` times) holds unmodified copies of the reference's files."""
    import hashlib
    import json
    from oracle import make_ref
    if not make_ref.staged():
        make_ref.make()
    man = json.load(open(os.path.join(make_ref.DST, "MANIFEST.json")))
    for rel, digest in man["files"].items():
        staged = open(os.path.join(make_ref.DST, "src", rel), "rb").read()
        assert hashlib.sha256(staged).hexdigest() == digest
        assert staged == open(os.path.join(make_ref.REF, rel), "rb").read(), rel
