"""GPU: randomized parity sweep.  The kernels replace IEEE divisions by reciprocal multiplications with exact
fall-backs, jump over constant TSDF regions and pick between two Stage A kernels; this sweep draws random grids,
voxel sizes, origins, intrinsics, poses (including cameras outside the grid and looking away), channel counts and
thresholds and checks the contract on each draw: indices / masks / counts / sums bit-exact, sample positions and
kept sets bit-exact, weights within 1e-5."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _random_scene(rng):
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


@pytest.mark.parametrize("seed", range(int(os.environ.get("CNRMA_RANDOM_SEEDS", "24"))))
def test_random_scene(seed):
    import cnrma_b200 as cn
    rng = np.random.default_rng(1000 + seed)
    s = _random_scene(rng)
    V, C, H, W = s["feats"].shape
    f = torch.from_numpy(s["feats"]).cuda().unsqueeze(1)
    if seed % 2:
        f = f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    p = torch.from_numpy(s["projs"]).cuda().unsqueeze(1)
    t = torch.from_numpy(s["tsdf"]).cuda()[None, None]
    args = (s["dim"], s["vs"], s["origin"], s["stride"])
    # Stage A
    px, py, valid = cn.project_views(p, *args, H, W)
    for v in range(V):
        ps = oracle.scale_projection(s["projs"][v], s["stride"])
        opx, opy, ovalid = oracle.project(s["dim"], s["vs"], s["origin"], ps, H, W)
        gv = valid[v, 0].cpu().numpy()
        assert np.array_equal(gv, ovalid)
        assert np.array_equal(px[v, 0].cpu().numpy()[gv], opx[ovalid]) and np.array_equal(py[v, 0].cpu().numpy()[gv], opy[ovalid])
    ovol, ocnt = oracle.aggregate_views(s["projs"], s["feats"], *args)
    vol, cnt, _ = cn.aggregate_views(p, f, *args)
    assert np.array_equal(cnt[0, 0].cpu().numpy(), ocnt)
    assert np.array_equal(vol[0].cpu().numpy().view(np.uint32), ovol.view(np.uint32))
    # Stage B
    rows = cn.rma_points(p, f, t, *args, grids=s["N"], threshold=s["thr"], normalize=False)[0].cpu().numpy()
    ref = oracle.aggregate_2d_features_ray_marching(s["projs"], s["feats"], s["tsdf"], *args, grids=s["N"],
                                                    neus_threshold=s["thr"], normalize=False)
    if ref is None:
        assert rows.shape[0] == 0
        return
    if rows.shape != ref.shape:      # a kept-set difference is only legitimate inside the threshold band
        near = np.abs(ref[:, 3] - np.float32(s["thr"])) <= 1e-5 * s["thr"]
        assert abs(rows.shape[0] - ref.shape[0]) <= int(near.sum()), (rows.shape, ref.shape)
        return
    assert np.array_equal(rows[:, :3].view(np.uint32), ref[:, :3].view(np.uint32))
    assert np.array_equal(rows[:, 4:].view(np.uint32), ref[:, 4:].view(np.uint32))
    assert np.abs(rows[:, 3] - ref[:, 3]).max() <= 1e-5 * max(ref[:, 3].max(), 1e-3)
