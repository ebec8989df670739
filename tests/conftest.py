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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_tree(npz):
    """Trees are stored as JSON with stringified keys; channel ids go back to int."""
    raw = json.loads(str(npz["tree"]))
    return {(int(k) if k != "None" else k): v for k, v in raw.items()}


def as_float_pairs(results):
    a = np.full(len(results), np.nan)
    b = np.full(len(results), np.nan)
    for i, r in enumerate(results):
        if isinstance(r, tuple):
            a[i], b[i] = float(r[0]), float(r[1])
        elif r is not None:
            a[i] = float(r)
    return a, b


def assert_same(got, want, rtol=0.0, what=""):
    got = np.asarray(got, dtype=float)
    want = np.asarray(want, dtype=float)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert (nan_g == nan_w).all(), f"{what}: NaN pattern differs at {np.flatnonzero(nan_g != nan_w)[:10]}"
    ok = ~nan_w
    if rtol == 0.0:
        bad = np.flatnonzero(got[ok] != want[ok])
    else:
        bad = np.flatnonzero(np.abs(got[ok] - want[ok]) > rtol * np.abs(want[ok]))
    assert len(bad) == 0, f"{what}: {len(bad)} mismatches, first idx {np.flatnonzero(ok)[bad[:5]]}: got {got[ok][bad[:5]]} want {want[ok][bad[:5]]}"


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
