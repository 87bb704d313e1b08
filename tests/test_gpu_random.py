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


from conftest import oracle_fusion, random_fusion_case  # noqa: E402
from conftest import random_scene as _random_scene  # noqa: E402


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


@pytest.mark.parametrize("seed", range(int(os.environ.get("CNRMA_RANDOM_SEEDS", "24"))))
@pytest.mark.parametrize("slab", [None, "1", "5", "cta", "pipe"])
def test_random_scene_stage_a_with_view_culling(seed, slab, monkeypatch):
    """The TMA gather kernel in column units: a conservative per-column cull of the views, then the exact projection over
    the survivors.  Random cameras (outside the grid, looking away, singular-ish intrinsics) and column lengths: a view
    that is culled wrongly would change a count or a sum."""
    import cnrma_b200 as cn
    monkeypatch.setenv("CNRMA_AGG_KERNEL", "tma")
    monkeypatch.setenv("CNRMA_AGG_CULL", {"cta": "2", "pipe": "0"}.get(slab, "1"))   # CTA / no cull (pipelined) / per warp
    if slab == "pipe":
        monkeypatch.setenv("CNRMA_AGG_PIPE", "1")
    if slab not in (None, "cta", "pipe"):
        monkeypatch.setenv("CNRMA_AGG_SLAB", slab)
    s = _random_scene(np.random.default_rng(3000 + seed))
    f = torch.from_numpy(s["feats"]).cuda().unsqueeze(1).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    p = torch.from_numpy(s["projs"]).cuda().unsqueeze(1)
    args = (s["dim"], s["vs"], s["origin"], s["stride"])
    for mean in (False, True):
        ovol, ocnt = oracle.aggregate_views(s["projs"], s["feats"], *args, mean=mean)
        vol, cnt, _ = cn.aggregate_views(p, f, *args, mean=mean)
        assert np.array_equal(cnt[0, 0].cpu().numpy(), ocnt)
        assert np.array_equal(vol[0].cpu().numpy().view(np.uint32), ovol.view(np.uint32))


@pytest.mark.parametrize("seed", range(int(os.environ.get("CNRMA_RANDOM_SEEDS", "24"))))
@pytest.mark.parametrize("jump", ["0", "1"])
def test_random_scene_march_jump_rules(seed, jump, monkeypatch):
    """Both jump rules of the NeuS march (clearance only / clearance + the sample's position inside its voxel) forced on
    every draw, whatever the launcher would pick: a jump that is one sample too long changes a kept set or a weight."""
    import cnrma_b200 as cn
    monkeypatch.setenv("CNRMA_MARCH_JUMP", jump)
    s = _random_scene(np.random.default_rng(5000 + seed))
    f = torch.from_numpy(s["feats"]).cuda().unsqueeze(1)
    p = torch.from_numpy(s["projs"]).cuda().unsqueeze(1)
    t = torch.from_numpy(s["tsdf"]).cuda()[None, None]
    args = (s["dim"], s["vs"], s["origin"], s["stride"])
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


@pytest.mark.parametrize("seed", range(8))
def test_random_fusion(seed):
    """GT TSDF fusion kernel against the oracle on random frame sets (half frame by frame, half in one launch)."""
    import cnrma_b200 as cn
    case = random_fusion_case(np.random.default_rng(9000 + seed))
    fus = cn.TSDFFusion(case["dim"], case["vs"], case["origin"], trunc_ratio=3, device="cuda", color=True, label=True)
    P, D = torch.from_numpy(case["P"]).cuda(), torch.from_numpy(case["depth"]).cuda()
    Cimg, L = torch.from_numpy(case["color"]).cuda(), torch.from_numpy(case["label"]).cuda()
    half = P.shape[0] // 2
    for i in range(half):
        fus.integrate(P[i], D[i], Cimg[i], L[i])
    if P.shape[0] > half:
        fus.integrate_frames(P[half:], D[half:], Cimg[half:], L[half:])
    tsdf, weight, col, lab = oracle_fusion(case)
    assert np.array_equal(fus.weight_vol.cpu().numpy().view(np.uint32), weight.view(np.uint32))
    assert np.array_equal(fus.tsdf_vol.cpu().numpy().view(np.uint32), tsdf.view(np.uint32))
    assert np.array_equal(fus.color_vol.cpu().numpy().view(np.uint32), col.view(np.uint32))
    assert np.array_equal(fus.label_vol.cpu().numpy(), lab)


@pytest.mark.parametrize("seed", range(8))
def test_random_head(seed):
    """TSDF head kernels against the oracle on random volume pyramids (NCDHW, channels-last, bf16), 1e-5; masks exact
    given the same previous scale."""
    import cnrma_b200 as cn
    rng = np.random.default_rng(11000 + seed)
    torch.manual_seed(11000 + seed)
    chans = [int(c) for c in rng.integers(2, 24, size=3)]
    coarse = tuple(int(v) for v in rng.integers(2, 7, size=3))
    thr = [float(t) for t in rng.choice([0.7, 0.9, 0.99], size=3)]
    prev = None
    for i, c in enumerate(chans[::-1]):
        dims = tuple(d * 2 ** i for d in coarse)
        x = float(rng.uniform(1.0, 4.0)) * torch.randn((1, c) + dims)
        w = torch.randn(c) / c ** 0.5
        for layout in ("ncdhw", "channels_last", "bf16"):
            xd = x.cuda()
            if layout == "channels_last":
                xd = xd.contiguous(memory_format=torch.channels_last_3d)
            if layout == "bf16":
                xd = xd.to(torch.bfloat16)
            t, m = cn.tsdf_head_scale(xd, w.cuda(), None if prev is None else torch.from_numpy(prev).cuda()[None, None], 1.05,
                                      thr[i - 1] if i > 0 else 0.0)
            ot, om = oracle.tsdf_head_scale(xd.float().cpu().numpy()[0], w.numpy(), prev, 1.05, thr[i - 1] if i > 0 else None)
            assert np.max(np.abs(t[0, 0].cpu().numpy() - ot)) <= 1e-5, (layout, i)
            assert np.array_equal(m[0, 0].cpu().numpy(), om), (layout, i)
        prev = ot
