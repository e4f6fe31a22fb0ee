import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def have_ref():
    return os.access(os.path.join(REFBIN, "bedops"), os.X_OK)


@pytest.fixture(scope="session")
def synth_files():
    """The seeded synthetic inputs named in tests/golden/synthetic.json -> {name: bytes}."""
    from bedops_b200 import synth
    spec = load_golden("synthetic.json")["files"]
    out = {}
    for name, a in spec.items():
        out[name] = synth.bed_text(a[0], a[1], tuple(a[2]), a[3], unique=len(a) > 4 and bool(a[4]),
                                   chroms=a[5] if len(a) > 5 else None)
    return out
