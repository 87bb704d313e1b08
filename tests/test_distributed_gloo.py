"""CPU, world_size 2 over gloo: the view-sharded partitioning / packing / all-reduce logic of
cnrma_b200.distributed, with oracle-backed compute hooks standing in for the CUDA kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import cnrma_b200 as cn
    from cnrma_b200 import distributed as D
    oracle.set_threads(1)
    sc = cn.synthetic.make_scene("small", seed=11)
    lo, hi = D.view_shard(sc.views, rank, world)

    def local_sums(projections, features, voxel_dim, voxel_size, origin, stride, packed):
        vol_view, cnt_view = packed
        s, c = oracle.aggregate_views(projections, features, voxel_dim, voxel_size, origin, stride, mean=False)
        vol_view[0].copy_(torch.from_numpy(s))
        cnt_view[0, 0].copy_(torch.from_numpy(c).float())

    def finalize(vol, cnt):
        return torch.where(cnt > 0, vol / cnt, torch.zeros(()))

    vol, cnt, valid = D.aggregate_views_sharded(sc.projections[lo:hi], sc.features[lo:hi], sc.voxel_dim, sc.voxel_size,
                                                sc.origin, sc.stride, local_sums=local_sums, finalize=finalize, batch=1,
                                                channels=sc.channels, device="cpu")

    def local_scatter(projections, features, tsdf, packed):
        wsum, wtot = packed
        rows = oracle.aggregate_2d_features_ray_marching(projections, features, tsdf, sc.voxel_dim, sc.voxel_size,
                                                         sc.origin, sc.stride, grids=sc.grids, normalize=False)
        s, t = oracle.dense_rma(rows, sc.voxel_dim, sc.voxel_size, sc.origin)
        wsum[0].copy_(torch.from_numpy(s))
        wtot[0, 0].copy_(torch.from_numpy(t))

    wsum, wtot = D.dense_rma_sharded(sc.projections[lo:hi], sc.features[lo:hi], sc.tsdf, sc.voxel_dim, sc.voxel_size,
                                     sc.origin, sc.stride, local_scatter=local_scatter, batch=1, channels=sc.channels,
                                     device="cpu")
    if rank == 0:
        ret["vol"], ret["cnt"], ret["valid"] = vol.numpy().copy(), cnt.numpy().copy(), valid.numpy().copy()
        ret["wsum"], ret["wtot"] = wsum.numpy().copy(), wtot.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_view_sharded_matches_single_process():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    import cnrma_b200 as cn
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        got = dict(ret)
    sc = cn.synthetic.make_scene("small", seed=11)
    ovol, ocnt = oracle.aggregate_views(sc.projections, sc.features, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    assert np.array_equal(got["cnt"][0, 0], ocnt.astype(np.float32))
    assert np.array_equal(got["valid"][0, 0], ocnt > 0)
    err = np.abs(got["vol"][0] - ovol).max() / np.abs(ovol).max()
    assert err <= 1e-5, err          # cross-rank sum order differs from view order
    rows = oracle.aggregate_2d_features_ray_marching(sc.projections, sc.features, sc.tsdf, sc.voxel_dim, sc.voxel_size,
                                                     sc.origin, sc.stride, grids=sc.grids, normalize=False)
    osum, otot = oracle.dense_rma(rows, sc.voxel_dim, sc.voxel_size, sc.origin)
    assert np.abs(got["wtot"][0, 0] - otot).max() <= 1e-5 * max(otot.max(), 1.0)
    assert np.abs(got["wsum"][0] - osum).max() <= 1e-5 * max(np.abs(osum).max(), 1.0)
