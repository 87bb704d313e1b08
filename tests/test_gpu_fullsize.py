"""GPU: BASELINE.json's full sizes, checked through size-independent properties of the path (the oracle
finishes in seconds only on a sample of the views; everything else is a property that must hold exactly):

  * Stage A count == number of per-view masks that see the voxel; valid == count > 0
  * folding the views in two chunks == one call (bit-exact); scaling features by a power of two scales the
    sums exactly (linearity)
  * channels-last and NCHW feature maps give identical results
  * Stage B: the row count equals the march's count, positions lie inside the grid, the normalised weights
    average to 1, rows of the un-normalised and normalised forms agree, per-view calls concatenate to the
    all-view call
  * the first views of the full-size scene match the oracle
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import assert_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cn():
    import cnrma_b200
    cnrma_b200.load()
    return cnrma_b200


@pytest.fixture(scope="module")
def big(cn):
    """cfg 2: 50 views x 256 ch @ 160x120, grid 80x80x32."""
    sc = cn.synthetic.make_scene("cfg2", seed=0, with_features=False)
    dev = torch.device("cuda")
    feats = cn.synthetic.device_features(sc, dev, channels_last=True)
    proj = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
    tsdf = torch.from_numpy(sc.tsdf).to(dev)[None, None]
    return sc, feats, proj, tsdf


def test_stage_a_count_matches_masks(cn, big):
    sc, feats, proj, _ = big
    vol, cnt, valid = cn.aggregate_views(proj, feats, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    _px, _py, masks = cn.project_views(proj, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, sc.height, sc.width)
    assert torch.equal(cnt.view(-1).long(), masks[:, 0].sum(0))
    assert torch.equal(valid, cnt > 0)
    assert bool(torch.isfinite(vol).all())
    assert float(vol[~valid.expand_as(vol)].abs().max()) == 0.0       # zero where no view sees the voxel


def test_stage_a_chunked_and_linear(cn, big):
    sc, feats, proj, _ = big
    full, cnt, _ = cn.aggregate_views(proj, feats, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, mean=False)
    out = cn.aggregate_views(proj[:17], feats[:17], sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, mean=False)
    two, cnt2, _ = cn.aggregate_views(proj[17:], feats[17:], sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                                      mean=False, out=out)
    assert torch.equal(cnt, cnt2)
    assert torch.equal(full.contiguous().view(torch.int32), two.contiguous().view(torch.int32))
    scaled, _, _ = cn.aggregate_views(proj, feats * 4.0, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, mean=False)
    assert torch.equal((full * 4.0).contiguous().view(torch.int32), scaled.contiguous().view(torch.int32))


def test_stage_a_layout_independent(cn, big):
    sc, feats, proj, _ = big
    a, ca, _ = cn.aggregate_views(proj[:8], feats[:8], sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    nchw = feats[:8].contiguous()                                      # reference layout -> transposer path
    assert nchw.stride(2) != 1
    b, cb, _ = cn.aggregate_views(proj[:8], nchw, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    assert torch.equal(ca, cb)
    assert torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))


def test_stage_a_first_views_match_oracle(cn, big):
    sc, feats, proj, _ = big
    v = 4
    f_host = feats[:v, 0].contiguous().cpu().numpy()
    ovol, ocnt = oracle.aggregate_views(sc.projections[:v], f_host, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    vol, cnt, _ = cn.aggregate_views(proj[:v], feats[:v], sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    assert np.array_equal(cnt[0, 0].cpu().numpy(), ocnt)
    assert np.array_equal(vol[0].cpu().numpy().view(np.uint32), ovol.view(np.uint32))


def test_stage_b_properties(cn, big):
    sc, feats, proj, tsdf = big
    v = 10
    (pts,), (st,) = cn.rma_points(proj[:v], feats[:v], tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                                  threshold=0.05, return_stats=True)
    (rows,) = cn.rma_points(proj[:v], feats[:v], tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                            threshold=0.05, normalize=False)
    assert pts.shape == (st["rows"], 3 + sc.channels) and rows.shape == (st["rows"], 4 + sc.channels)
    assert torch.equal(pts[:, :3], rows[:, :3])
    w = rows[:, 3]
    assert float(w.min()) >= np.float32(0.05)
    assert_rel(float(w.double().sum()), st["weight_sum"], 1e-9, what="weight sum")
    # kept samples lie inside the grid (rounded ids in range)
    ids = torch.round(pts[:, :3] / sc.voxel_size)
    dims = torch.tensor(sc.voxel_dim, device=ids.device, dtype=ids.dtype)
    assert bool(((ids >= 0) & (ids < dims)).all())
    # normalised rows == un-normalised rows * w / mean(w)
    wn = (w / torch.tensor(st["mean"], dtype=torch.float32, device=w.device)).unsqueeze(1)   # tensor / tensor: IEEE division
    assert torch.equal(pts[:, 3:], rows[:, 4:] * wn)
    assert abs(float((w.double() / st["mean"]).mean()) - 1.0) < 1e-6
    # per-view calls concatenate to the all-view call (rm.py:289-293)
    chunks = [cn.rma_points(proj[i:i + 1], feats[i:i + 1], tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                            threshold=0.05, normalize=False)[0] for i in range(3)]
    cat = torch.cat(chunks, dim=0)
    assert torch.equal(cat, rows[: cat.shape[0]])


def test_stage_b_first_view_matches_oracle(cn, big):
    sc, feats, proj, tsdf = big
    f_host = feats[:1, 0].contiguous().cpu().numpy()
    ref = oracle.aggregate_2d_features_ray_marching(sc.projections[:1], f_host, sc.tsdf, sc.voxel_dim, sc.voxel_size,
                                                    sc.origin, sc.stride, grids=sc.grids, normalize=False)
    (rows,) = cn.rma_points(proj[:1], feats[:1], tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                            threshold=0.05, normalize=False)
    got = rows.cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got[:, :3].view(np.uint32), ref[:, :3].view(np.uint32))
    assert np.array_equal(got[:, 4:].view(np.uint32), ref[:, 4:].view(np.uint32))
    assert_rel(got[:, 3], ref[:, 3], 1e-5, what="weights")


def test_dense_rma_totals(cn, big):
    sc, feats, proj, tsdf = big
    v = 6
    wsum, wtot = cn.dense_rma(proj[:v], feats[:v], tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                              threshold=0.05)
    (rows,), (st,) = cn.rma_points(proj[:v], feats[:v], tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                                   threshold=0.05, normalize=False, return_stats=True)
    assert_rel(float(wtot.double().sum()), st["weight_sum"], 1e-5, what="total weight")
    ref = (rows[:, 3:4].double() * rows[:, 4:].double()).sum(0)
    got = wsum[0].double().sum(dim=(1, 2, 3))
    assert float((got - ref).abs().max()) <= 1e-5 * float((rows[:, 3:4] * rows[:, 4:]).abs().double().sum(0).max())


@pytest.mark.parametrize("name", ["cfg3", "cfg4_c32"])
def test_other_configs_run(cn, name):
    """ARKit-shaped (cfg 3) and fine-grid (cfg 4, C=32) sizes: the invariants at those shapes, a few views."""
    sc = cn.synthetic.make_scene(name, seed=1, with_features=False)
    dev = torch.device("cuda")
    v = 6
    sc.projections = sc.projections[:v]
    feats = cn.synthetic.device_features(sc, dev, channels_last=True)
    proj = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
    tsdf = torch.from_numpy(sc.tsdf).to(dev)[None, None]
    vol, cnt, valid = cn.aggregate_views(proj, feats, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    _px, _py, masks = cn.project_views(proj, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, sc.height, sc.width)
    assert torch.equal(cnt.view(-1).long(), masks[:, 0].sum(0))
    (pts,), (st,) = cn.rma_points(proj, feats, tsdf, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, threshold=0.05,
                                  return_stats=True)
    assert pts.shape[0] == st["rows"] > 0
    assert bool(torch.isfinite(pts).all())
