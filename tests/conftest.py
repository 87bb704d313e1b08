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


@pytest.fixture(autouse=True)
def _tuning_knobs_follow_the_environment(monkeypatch):
    """The library reads its CNRMA_* knobs once per process.  Tests flip them with monkeypatch.setenv / os.environ: this
    wraps setenv / delenv so that the library re-reads them right away, and re-reads once more after the test (when
    monkeypatch has restored the environment)."""
    import cnrma_b200
    loaded = os.path.exists(cnrma_b200.LIB_PATH)
    if loaded:
        real_set, real_del = monkeypatch.setenv, monkeypatch.delenv

        def setenv(name, value, *a, **kw):
            real_set(name, value, *a, **kw)
            if name.startswith("CNRMA_"):
                cnrma_b200.reload_tuning()

        def delenv(name, *a, **kw):
            real_del(name, *a, **kw)
            if name.startswith("CNRMA_"):
                cnrma_b200.reload_tuning()

        monkeypatch.setenv, monkeypatch.delenv = setenv, delenv
    yield
    if loaded:
        monkeypatch.undo()
        cnrma_b200.reload_tuning()


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


def random_scene(rng):
    """A random small scene: grid, voxel size, origin, intrinsics, poses (also outside the grid / looking away), channel
    count, one of three TSDF families, step count and threshold (shared by the GPU sweep and the oracle-vs-reference sweep)."""
    dim = tuple(int(v) for v in rng.integers(3, 28, size=3))
    vs = float(rng.choice([0.04, 0.08, 0.1, 0.137, 0.25, 0.3333]))
    origin = (rng.uniform(-1, 1, size=3) * rng.choice([0.0, 1.0])).astype(np.float32)
    V = int(rng.integers(1, 9))
    C = int(rng.choice([4, 8, 12, 16, 32, 64]))
    H, W = int(rng.integers(6, 40)), int(rng.integers(6, 48))
    stride = int(rng.choice([1, 2, 4]))
    extent = np.array(dim) * vs
    projs = np.empty((V, 3, 4), np.float32)
    for v in range(V):
        f = rng.uniform(0.5, 1.5) * W * stride
        k = np.array([[f, 0, rng.uniform(0.3, 0.7) * W * stride], [0, f * rng.uniform(0.9, 1.1), rng.uniform(0.3, 0.7) * H * stride],
                      [0, 0, 1.0]])
        cam = origin + extent * rng.uniform(-0.3, 1.3, size=3)              # sometimes outside the grid
        target = origin + extent * rng.uniform(0.0, 1.0, size=3)
        fwd = target - cam
        fwd /= max(np.linalg.norm(fwd), 1e-9)
        if rng.random() < 0.15:
            fwd = -fwd                                                      # looking away
        up = np.array([0.0, 0.0, 1.0]) if abs(fwd[2]) < 0.95 else np.array([0.0, 1.0, 0.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        pose = np.eye(4)
        pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, down, fwd, cam
        projs[v] = (k @ np.linalg.inv(pose)[:3]).astype(np.float32)
    feats = rng.standard_normal((V, C, H, W), dtype=np.float32)
    kind = rng.integers(0, 3)
    if kind == 0:
        tsdf = rng.uniform(-1, 1, size=dim).astype(np.float32)
    elif kind == 1:                                                          # piecewise constant with a ramp: big empty regions
        g = np.indices(dim).astype(np.float32)
        d = np.minimum.reduce([g[a] - 0.2 * dim[a] for a in range(3)] + [0.8 * dim[a] - g[a] for a in range(3)])
        tsdf = (np.clip(-d / 2.0, -1, 1) * 0.999).astype(np.float32)
    else:                                                                    # blocks of constant value
        coarse = rng.choice([-0.999, 0.999, 0.3], size=[(n + 3) // 4 for n in dim]).astype(np.float32)
        tsdf = np.kron(coarse, np.ones((4, 4, 4), np.float32))[: dim[0], : dim[1], : dim[2]].copy()
    N = int(rng.choice([40, 97, 300]))
    thr = float(rng.choice([0.05, 0.02, 0.3]))
    return dict(dim=dim, vs=vs, origin=origin, stride=stride, projs=projs, feats=feats, tsdf=tsdf, N=N, thr=thr)


def random_fusion_case(rng):
    """Random grid / origin / frames / depth maps with holes for the GT TSDF fusion sweeps."""
    import cnrma_b200
    dim = tuple(int(v) for v in rng.integers(6, 26, size=3))
    vs = float(rng.choice([0.1, 0.25, 0.3]))
    origin = tuple(float(x) for x in (rng.uniform(-0.5, 0.5, size=3) * rng.choice([0.0, 1.0])))
    frames, h, w = int(rng.integers(1, 7)), int(rng.integers(12, 40)), int(rng.integers(16, 48))
    extent = tuple(d * vs for d in dim)
    P, k, poses = cnrma_b200.synthetic.ring_cameras(frames, h, w, 1, extent, rng, return_poses=True)
    depth = cnrma_b200.synthetic.room_depth_maps(k, poses, h, w, extent, rng)
    depth[rng.random(depth.shape) < 0.1] = 0.0                     # holes: no reading
    color = rng.uniform(0, 255, size=(frames, 3, h, w)).astype(np.float32)
    label = rng.integers(0, 40, size=(frames, h, w)).astype(np.int64)
    return dict(dim=dim, vs=vs, origin=origin, P=P, depth=depth, color=color, label=label)


def oracle_fusion(case):
    """The oracle's volumes for a random_fusion_case, frame by frame."""
    import oracle
    n = int(np.prod(case["dim"]))
    tsdf, weight = np.ones(n, np.float32), np.zeros(n, np.float32)
    col, lab = np.zeros((3, n), np.float32), -np.ones(n, np.int64)
    for i in range(case["P"].shape[0]):
        oracle.tsdf_integrate(case["dim"], case["vs"], np.float32(case["origin"]), case["P"][i], case["depth"][i],
                              case["vs"] * 3, tsdf, weight, case["color"][i], col, case["label"][i], lab)
    return tsdf, weight, col, lab
