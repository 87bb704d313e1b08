import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# oracle/ is test infrastructure: importable from tests only
if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    g["voxel_dim"] = tuple(int(x) for x in g["voxel_dim"])
    g["voxel_size"] = float(g["voxel_size"])
    g["stride"] = int(g["stride"])
    g["grids"] = int(g["grids"])
    g["thr"] = float(g["thr"])
    return g


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)


def assert_rel(a, b, tol, floor=1e-6, what=""):
    """max |a-b| / max(|b|, floor) <= tol"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return
    err = np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))
    assert err <= tol, f"{what}: max relative error {err:.3e} > {tol:.1e}"
