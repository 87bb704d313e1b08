"""Point-cloud hand-off (rm.py:339-407): host mask logic on CPU, device selection on GPU, against the reference's
own switch_pointcloud output (tests/golden)."""
import numpy as np
import pytest
import torch

import cnrma_b200 as cn


def test_sample_points_mask_matches_reference_rng(golden):
    g = golden
    n = g["neus_points"].shape[0]
    np.random.seed(2024)
    mask = cn.sample_points(n, int(g["handoff_max_points"]))
    assert mask.dtype == bool and mask.sum() == int(g["handoff_max_points"])
    assert np.array_equal(mask, g["handoff_mask"].astype(bool))
    assert cn.sample_points(5, 10).all()                       # fewer points than max_points: keep everything


@pytest.mark.gpu
def test_switch_pointcloud_bit_exact(golden):
    g = golden
    pts = torch.from_numpy(g["neus_points"]).cuda()
    coords, feats = cn.switch_pointcloud([pts], [g["handoff_offset"]], masks=[g["handoff_mask"].astype(bool)])
    assert np.array_equal(coords[0].cpu().numpy().view(np.uint32), g["handoff_coords"].view(np.uint32))
    assert np.array_equal(feats[0].cpu().numpy().view(np.uint32), g["handoff_features"].view(np.uint32))
    # drawing the mask with the reference's seed gives the same rows
    np.random.seed(2024)
    coords2, feats2 = cn.switch_pointcloud([pts], [g["handoff_offset"]], max_points=int(g["handoff_max_points"]))
    assert torch.equal(coords2[0], coords[0]) and torch.equal(feats2[0], feats[0])
    # no sampling: only the offset is applied
    c3, f3 = cn.switch_pointcloud([pts], [g["handoff_offset"]])
    assert torch.equal(f3[0], pts[:, 3:])
    assert torch.equal(c3[0], pts[:, :3] + torch.from_numpy(g["handoff_offset"]).cuda())


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["packed", "scalar"])
def test_fused_selected_fill_equals_two_steps(golden, kernel, monkeypatch):
    """march + fill of only the kept rows == full point cloud followed by switch_pointcloud (both forms of the selecting
    fill: the packed kernel's SELECT variant and the plain scalar kernel)."""
    if kernel == "scalar":
        monkeypatch.setenv("CNRMA_FILL_SELECT_KERNEL", "scalar")
    g = golden
    f = torch.from_numpy(g["features"]).cuda().unsqueeze(1)
    p = torch.from_numpy(g["projections"]).cuda().unsqueeze(1)
    t = torch.from_numpy(g["tsdf"]).cuda()[None, None]
    args = (g["voxel_dim"], g["voxel_size"], g["origin"], g["stride"])
    full = cn.rma_points(p, f, t, *args, grids=g["grids"], threshold=g["thr"])[0]
    n = full.shape[0]
    rng = np.random.RandomState(5)
    mask = cn.sample_points(n, max(1, n // 4), rng)
    want_c, want_f = cn.switch_pointcloud([full], [g["handoff_offset"]], masks=[mask])
    got_c, got_f = cn.rma_points_selected(p, f, t, *args, offsets=[g["handoff_offset"]], masks=[mask],
                                          grids=g["grids"], threshold=g["thr"])
    assert torch.equal(got_c[0], want_c[0]) and torch.equal(got_f[0], want_f[0])
    if full.shape == g["neus_points"].shape:
        m2 = g["handoff_mask"].astype(bool)
        got_c, got_f = cn.rma_points_selected(p, f, t, *args, offsets=[g["handoff_offset"]], masks=[m2],
                                              grids=g["grids"], threshold=g["thr"])
        assert np.array_equal(got_c[0].cpu().numpy().view(np.uint32), g["handoff_coords"].view(np.uint32))
        err = np.abs(got_f[0].cpu().numpy() - g["handoff_features"]).max() / np.abs(g["handoff_features"]).max()
        assert err <= 1e-5


@pytest.mark.gpu
def test_mirror_switch_pointcloud(golden):
    g = golden
    ag = cn.RayMarchingAggregator(g["voxel_size"], g["voxel_dim"], origin=g["origin"].tolist(), neus_threshold=g["thr"],
                                  max_points=int(g["handoff_max_points"]))
    pts = torch.from_numpy(g["neus_points"]).cuda()
    np.random.seed(2024)
    coords, feats, boxes = ag.switch_pointcloud([pts], ["boxes"], [torch.from_numpy(g["handoff_offset"])], True)
    assert boxes == ["boxes"]
    assert np.array_equal(coords[0].cpu().numpy().view(np.uint32), g["handoff_coords"].view(np.uint32))
    assert np.array_equal(feats[0].cpu().numpy().view(np.uint32), g["handoff_features"].view(np.uint32))


@pytest.mark.gpu
def test_device_sampler_properties():
    """The on-device alternative to sample_points: exact count, deterministic per seed, uniform, usable as a mask."""
    n, k = 200_000, 30_000
    m1 = cn.sample_points_device(n, k, seed=7, device="cuda")
    m2 = cn.sample_points_device(n, k, seed=7, device="cuda")
    m3 = cn.sample_points_device(n, k, seed=8, device="cuda")
    assert m1.dtype == torch.bool and int(m1.sum()) == k
    assert torch.equal(m1, m2) and not torch.equal(m1, m3) and int(m3.sum()) == k
    # uniformity: counts in 100 equal slices of the index range follow Binomial(2000, 0.15): std ~16
    per_slice = m1.view(100, -1).sum(1).float()
    assert abs(float(per_slice.mean()) - k / 100) < 1e-3
    assert float(per_slice.std()) < 3 * (2000 * 0.15 * 0.85) ** 0.5
    # overlap of two independent draws ~ k*k/n
    both = int((m1 & m3).sum())
    assert abs(both - k * k / n) < 6 * (k * k / n) ** 0.5
    assert bool(cn.sample_points_device(10, 25, seed=1, device="cuda").all())
    assert not bool(cn.sample_points_device(10, 0, seed=1, device="cuda").any())
    for kk in (1, 9, 10):
        assert int(cn.sample_points_device(10, kk, seed=3, device="cuda").sum()) == min(kk, 10)
    pts = torch.randn(n, 7, device="cuda")
    coords, feats = cn.switch_pointcloud([pts], [[0.0, 0.0, 0.0]], masks=[m1])
    assert torch.equal(torch.cat((coords[0], feats[0]), 1), pts[m1])
