"""More views than one launch carries device pointers for (kMaxViewsPerLaunch = 512): Stage A, the point rows, the
dense form and both backward kernels split the view list over several launches; results must not depend on it."""
import numpy as np
import pytest
import torch

import oracle
from conftest import assert_rel

pytestmark = pytest.mark.gpu
VIEWS = 530


@pytest.fixture(scope="module")
def cn():
    import cnrma_b200
    cnrma_b200.load()
    return cnrma_b200


@pytest.fixture(scope="module")
def wide(cn):
    # 128 fp32 channels = 512-byte rows: the TMA Stage A kernel (two launches: 512 + 18 views);
    # 6 x 8 pixels per view, so the fill's view groups must be multiples of 16 views (256-ray blocks)
    sc = cn.synthetic.make_scene(dict(views=VIEWS, channels=128, height=6, width=8, voxel_dim=(10, 10, 4),
                                      voxel_size=0.5, tsdf="room", grids=40, dtype="f32"), seed=5)
    p = torch.from_numpy(sc.projections).cuda().unsqueeze(1)
    f = torch.from_numpy(sc.features).cuda().unsqueeze(1).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    t = torch.from_numpy(sc.tsdf).cuda()[None, None]
    return sc, p, f, t


def test_stage_a_over_512_views(cn, wide):
    sc, p, f, _ = wide
    ovol, ocnt = oracle.aggregate_views(sc.projections, sc.features, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    vol, cnt, _ = cn.aggregate_views(p, f, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    assert int(ocnt.max()) > 64
    assert np.array_equal(cnt[0, 0].cpu().numpy(), ocnt)
    assert np.array_equal(vol[0].cpu().numpy().view(np.uint32), ovol.view(np.uint32))
    px, py, valid = cn.project_views(p, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, sc.height, sc.width)
    assert np.array_equal(valid[:, 0].sum(0).cpu().numpy().reshape(ocnt.shape), ocnt)


def test_points_over_512_views(cn, wide):
    sc, p, f, t = wide
    ref = oracle.aggregate_2d_features_ray_marching(sc.projections, sc.features, sc.tsdf, sc.voxel_dim, sc.voxel_size,
                                                    sc.origin, sc.stride, grids=sc.grids, neus_threshold=0.05)
    pts = cn.rma_points(p, f, t, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, grids=sc.grids, threshold=0.05)[0]
    pts = pts.cpu().numpy()
    assert pts.shape == ref.shape and ref.shape[0] > 1000
    assert np.array_equal(pts[:, :3].view(np.uint32), ref[:, :3].view(np.uint32))
    assert_rel(pts[:, 3:], ref[:, 3:], 1e-5, what="points")
    wsum, wtot = cn.dense_rma(p, f, t, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, grids=sc.grids, threshold=0.05)
    rows = oracle.aggregate_2d_features_ray_marching(sc.projections, sc.features, sc.tsdf, sc.voxel_dim, sc.voxel_size,
                                                     sc.origin, sc.stride, grids=sc.grids, normalize=False)
    _osum, otot = oracle.dense_rma(rows, sc.voxel_dim, sc.voxel_size, sc.origin)
    assert abs(float(wtot.sum()) - float(otot.sum())) <= 1e-4 * float(otot.sum())


def test_backward_over_512_views_matches_split_runs(cn, wide):
    """Gradients of the un-averaged / un-normalised lifts are independent per view, so the 530-view backward must equal
    the backward of the two halves run separately (each within one launch)."""
    sc, p, f, t = wide
    args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    h = VIEWS // 2

    def grads(pp, ff):
        ff = ff.detach().clone().requires_grad_(True)
        vol, _c, _v = cn.aggregate_views(pp, ff, *args, mean=False)
        torch.manual_seed(1)
        ga, = torch.autograd.grad(vol, ff, torch.randn(vol.shape, device="cuda"))
        rows = cn.rma_points(pp, ff, t, *args, grids=sc.grids, threshold=0.05, normalize=False)[0]
        gb, = torch.autograd.grad(rows, ff, torch.ones_like(rows))
        return ga, gb

    ga, gb = grads(p, f)
    ga0, gb0 = grads(p[:h], f[:h])
    ga1, gb1 = grads(p[h:], f[h:])
    ref_a, ref_b = torch.cat((ga0, ga1)), torch.cat((gb0, gb1))
    assert float(ref_a.abs().max()) > 0 and float(ref_b.abs().max()) > 0
    assert float((ga - ref_a).abs().max()) <= 1e-5 * float(ref_a.abs().max())      # fp32 reductions in the memory system
    assert torch.equal(gb, ref_b)                                                    # per-ray sums: no atomics


@pytest.mark.parametrize("channels,dtype", [(16, torch.float32), (32, torch.float32), (48, torch.float32), (64, torch.bfloat16),
                                            (96, torch.float32)])
def test_stage_a_long_lists_with_segments(cn, channels, dtype):
    """Short rows and hundreds of views: the long-list Stage A kernel (all views of a voxel in one go).  Every camera
    looks at the same small grid, so many voxels are seen by more views than a list holds (161) and are served in
    segments through the volume; also as two accumulating calls, and on a box.  Bit-exact against the oracle."""
    views = 400
    sc = cn.synthetic.make_scene(dict(views=views, channels=channels, height=6, width=8, voxel_dim=(11, 9, 5),
                                      voxel_size=0.5, tsdf="room", grids=40, dtype="f32"), seed=8)
    p = torch.from_numpy(sc.projections).cuda().unsqueeze(1)
    f = torch.from_numpy(sc.features).cuda().to(dtype).unsqueeze(1).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    feats_host = f[:, 0].float().contiguous().cpu().numpy()
    args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    for mean in (True, False):
        ovol, ocnt = oracle.aggregate_views(sc.projections, feats_host, *args, mean=mean)
        assert int(ocnt.max()) > 200
        vol, cnt, valid = cn.aggregate_views(p, f, *args, mean=mean)
        assert np.array_equal(cnt[0, 0].cpu().numpy(), ocnt)
        assert np.array_equal(valid[0, 0].cpu().numpy(), ocnt > 0)
        assert np.array_equal(vol[0].cpu().numpy().view(np.uint32), ovol.view(np.uint32)), (channels, mean)
    # two accumulating calls (the second starts from sums and counts in the volume), then the mean
    out = cn.aggregate_views(p[:170], f[:170], *args, mean=False)
    vol2, cnt2, _ = cn.aggregate_views(p[170:], f[170:], *args, mean=True, out=out)
    ovol, ocnt = oracle.aggregate_views(sc.projections, feats_host, *args, mean=True)
    assert np.array_equal(cnt2[0, 0].cpu().numpy(), ocnt)
    assert np.array_equal(vol2[0].cpu().numpy().view(np.uint32), ovol.view(np.uint32))
    # a box of the grid
    lo, dim = (2, 1, 1), (7, 6, 3)
    bv, bc, _ = cn.aggregate_views(p, f, *args, mean=True, box=(lo, dim))
    sl = tuple(slice(l, l + d) for l, d in zip(lo, dim))
    assert np.array_equal(bc[0, 0].cpu().numpy(), ocnt[sl])
    assert np.array_equal(bv[0].contiguous().cpu().numpy().view(np.uint32), np.ascontiguousarray(ovol[(slice(None),) + sl]).view(np.uint32))
