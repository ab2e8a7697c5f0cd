import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def pins():
    with open(os.path.join(GOLDEN, "pins.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def pins_large():
    """Reference pins at presets L and paper (tests/golden/make_golden_large.py)."""
    with open(os.path.join(GOLDEN, "pins_large.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cases():
    """tests/golden/cases.npz regrouped as {bench: [ {field: array}, ... ]}."""
    z = np.load(os.path.join(GOLDEN, "cases.npz"))
    out = {}
    for key in z.files:
        bench, idx, field = key.split(".", 2)
        out.setdefault(bench, {}).setdefault(int(idx), {})[field] = z[key]
    return {b: [d[i] for i in sorted(d)] for b, d in out.items()}


def assert_bit_equal(got, want, what=""):
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s != %s" % (what, got.shape, want.shape)
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        first = tuple(bad[0])
        with np.errstate(all="ignore"):
            rel = np.nanmax(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
        raise AssertionError("%s: %d / %d cells differ; first at %s got %r want %r; max rel err %.3e"
                             % (what, len(bad), got.size, first, got[first], want[first], rel))
