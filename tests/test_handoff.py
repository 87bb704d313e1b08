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
def test_fused_selected_fill_equals_two_steps(golden):
    """march + fill of only the kept rows == full point cloud followed by switch_pointcloud."""
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
