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


@pytest.mark.gpu
@pytest.mark.parametrize("n,c,scale", [(1, 4, 1.0), (5000, 8, 0.05), (200000, 32, 0.4), (70000, 259 - 3, 2.0)])
def test_quantize_points_keeps_the_first_row_of_every_cell(n, c, scale):
    """rm.py:330-332 + MinkowskiEngine's collate / unique: cells = trunc(coords / 0.01), first row of every cell in row
    order -- bit-exact against the numpy restatement (oracle.quantize_unique_first), duplicates, negative coordinates
    and cell boundaries included."""
    import cnrma_b200 as cn
    import oracle
    rng = np.random.default_rng(n)
    coords = (rng.standard_normal((n, 3)) * scale).astype(np.float32)
    coords[::7] = coords[::7].round(2)                        # values on cell boundaries
    if n > 10:
        coords[n // 2:] = coords[: n - n // 2] + rng.uniform(-0.004, 0.004, size=(n - n // 2, 3)).astype(np.float32)  # near-duplicates
        coords[3] = [-0.0, 0.0, -0.0099]
    feats = rng.standard_normal((n, c)).astype(np.float32)
    want_q, want_f, want_c = oracle.quantize_unique_first(coords, feats, 0.01)
    rows = torch.from_numpy(np.concatenate([coords, feats], 1)).cuda()
    for cd, fd in ((rows[:, :3], rows[:, 3:]),                                       # views of one row buffer
                   (torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda())):   # separate tensors
        q, f, cc = cn.quantize_points(cd, fd, 0.01)
        assert np.array_equal(q.cpu().numpy(), want_q)
        assert np.array_equal(f.cpu().numpy().view(np.uint32), want_f.view(np.uint32))
        assert np.array_equal(cc.cpu().numpy().view(np.uint32), want_c.view(np.uint32))
    assert want_q.shape[0] < n or n <= 10
    coords4, feats_cat = cn.sparse_collate_quantized([rows[:, :3], rows[:, :3] + 1.0], [rows[:, 3:], rows[:, 3:]], 0.01)
    assert coords4.shape[1] == 4 and int(coords4[:, 0].max()) == 1 and coords4.shape[0] == feats_cat.shape[0]
    assert np.array_equal(coords4[: want_q.shape[0], 1:].cpu().numpy(), want_q)


@pytest.mark.gpu
def test_quantize_points_on_the_golden_hand_off(golden):
    """The reference's own switch_pointcloud output (golden vectors), quantised: device == numpy restatement."""
    import cnrma_b200 as cn
    import oracle
    coords, feats = golden["handoff_coords"], golden["handoff_features"]
    want_q, want_f, _ = oracle.quantize_unique_first(coords, feats, 0.01)
    q, f, _ = cn.quantize_points(torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda(), 0.01)
    assert np.array_equal(q.cpu().numpy(), want_q) and np.array_equal(f.cpu().numpy(), want_f)


def test_quantize_restatement_properties():
    """CPU: the numpy restatement itself -- truncation toward zero, first occurrence wins, order kept."""
    import oracle
    coords = np.array([[0.019, -0.019, 0.0], [0.011, -0.011, 0.009], [0.5, 0.5, 0.5], [-0.005, 0.005, 0.0],
                       [0.5001, 0.5, 0.5]], np.float32)
    feats = np.arange(5, dtype=np.float32)[:, None]
    q, f, c = oracle.quantize_unique_first(coords, feats, 0.01)
    assert q.tolist() == [[1, -1, 0], [50, 50, 50], [0, 0, 0]]      # (-0.5 .. 0.5) truncates to 0; rows 1 and 4 are duplicates
    assert f[:, 0].tolist() == [0.0, 2.0, 3.0]
