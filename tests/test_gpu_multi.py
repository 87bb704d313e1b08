"""GPU, >= 2 devices (skipped otherwise): the view-sharded paths over real NCCL / NVLink against the single-GPU result,
at world sizes 2, 4 and 8 (whatever the box has).
Run with `gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import cnrma_b200 as cn
    from cnrma_b200 import distributed as D
    sc = cn.synthetic.make_scene("small", seed=21, views=11)
    lo, hi = D.view_shard(sc.views, rank, world)
    f = torch.from_numpy(sc.features).to(dev).unsqueeze(1).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    p = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
    t = torch.from_numpy(sc.tsdf).to(dev)[None, None]
    args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    vol, cnt, valid = D.aggregate_views_sharded(p[lo:hi], f[lo:hi], *args)
    wsum, wtot = D.dense_rma_sharded(p[lo:hi], f[lo:hi], t, *args, grids=sc.grids, threshold=0.05)
    pts = D.rma_points_sharded(p[lo:hi], f[lo:hi], t, *args, grids=sc.grids, threshold=0.05)[0]
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([pts.shape[0]], dtype=torch.int64, device=dev))
    if rank == 0:
        # single-GPU results on the same device
        v1, c1, _ = cn.aggregate_views(p, f, *args)
        s1, t1 = cn.dense_rma(p, f, t, *args, grids=sc.grids, threshold=0.05)
        p1 = cn.rma_points(p, f, t, *args, grids=sc.grids, threshold=0.05)[0]
        ret["cnt_equal"] = bool(torch.equal(cnt.view(-1), c1.view(-1).float()))
        ret["vol_err"] = float((vol - v1).abs().max() / v1.abs().max())
        ret["wtot_err"] = float((wtot - t1).abs().max() / t1.abs().max())
        ret["wsum_err"] = float((wsum - s1).abs().max() / s1.abs().max())
        n0 = int(sizes[0])
        ret["rows_total"] = int(sum(int(s) for s in sizes)) == p1.shape[0]
        ret["pts_err"] = float((pts - p1[:n0]).abs().max() / p1.abs().max())
    dist.barrier()
    dist.destroy_process_group()


WORLDS = [2, 4, 8]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", WORLDS)
def test_view_sharded_over_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() % 1000) + world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        got = dict(ret)
    assert got["cnt_equal"]
    assert got["vol_err"] <= 1e-5
    assert got["wtot_err"] <= 1e-5 and got["wsum_err"] <= 1e-5
    assert got["rows_total"]
    assert got["pts_err"] <= 1e-5


def _worker_p2p(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import cnrma_b200 as cn
    from cnrma_b200 import distributed as D
    ok = True
    for channels in (16, 256):                       # list kernel / TMA kernel
        sc = cn.synthetic.make_scene("small", seed=22, channels=channels, views=7)   # world 8: one rank has no view
        lo, hi = D.view_shard(sc.views, rank, world)
        f = torch.from_numpy(sc.features).to(dev).unsqueeze(1).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        p = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
        args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
        v1, c1, m1 = cn.aggregate_views(p, f, *args)                      # all views on this GPU
        v1 = v1[0].permute(1, 2, 3, 0).reshape(-1, channels)
        for _ in range(3):                                                # repeated calls reuse the symmetric buffer
            a, b, vol, cnt, valid = D.aggregate_views_p2p(p[lo:hi], f[lo:hi], *args, channels=channels, device=dev)
        ok = ok and bool(torch.equal(cnt, c1.view(-1)[a:b])) and bool(torch.equal(valid, m1.view(-1)[a:b]))
        err = float((vol - v1[a:b]).abs().max() / v1.abs().max())
        ok = ok and err <= 1e-5 and (b - a) > 0
    ret[f"ok{rank}"] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", WORLDS)
def test_view_sharded_over_peer_memory(world):
    """aggregate_views_p2p: the kernel's own NVLink stores into symmetric memory instead of an all-reduce."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + (os.getpid() % 1000) + world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_p2p, args=(world, port, ret), nprocs=world, join=True)
        got = dict(ret)
    assert all(got.get(f"ok{r}") for r in range(world)), got


def _worker_pipelined(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import cnrma_b200 as cn
    from cnrma_b200 import distributed as D
    ok, why = True, ""
    for channels in (16, 256):
        sc = cn.synthetic.make_scene("small", seed=23, channels=channels, views=11)
        lo, hi = D.view_shard(sc.views, rank, world)
        f = torch.from_numpy(sc.features).to(dev).unsqueeze(1).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        p = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
        args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
        v1, c1, m1 = cn.aggregate_views(p, f, *args)
        scale = float(v1.abs().max())
        for chunks in (1, 3):
            vol, cnt, valid = D.aggregate_views_sharded(p[lo:hi], f[lo:hi], *args, chunks=max(chunks, 2))
            good = torch.equal(cnt.view(-1), c1.view(-1).float()) and torch.equal(valid, m1) and \
                float((vol - v1).abs().max()) / scale <= 1e-5
            ok, why = ok and good, why or ("" if good else f"all_reduce chunks={chunks} C={channels}")
            x0, x1, vol, cnt, valid = D.aggregate_views_sharded(p[lo:hi], f[lo:hi], *args, chunks=chunks,
                                                                collective="reduce_scatter")
            good = torch.equal(cnt[0, 0], c1[0, 0, x0:x1].float()) and torch.equal(valid[0, 0], m1[0, 0, x0:x1]) and \
                float((vol[0] - v1[0, :, x0:x1]).abs().max()) / scale <= 1e-5 and x1 > x0
            ok, why = ok and good, why or ("" if good else f"reduce_scatter chunks={chunks} C={channels}")
    ret[f"ok{rank}"] = ok
    ret[f"why{rank}"] = why
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", WORLDS)
def test_view_sharded_pipelined_collectives(world):
    """Chunk-pipelined all-reduce and reduce-to-owner (reduce-scatter) forms of the view-sharded Stage A."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29800 + (os.getpid() % 1000) + world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_pipelined, args=(world, port, ret), nprocs=world, join=True)
        got = dict(ret)
    assert all(got.get(f"ok{r}") for r in range(world)), got


def _worker_exchange(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import cnrma_b200 as cn
    from cnrma_b200 import distributed as D
    ok, why = True, ""
    for channels, views, dtype in ((16, 11, torch.float32), (256, 9, torch.float32), (128, 5, torch.bfloat16)):
        sc = cn.synthetic.make_scene("small", seed=24, channels=channels, views=views)
        lo, hi = D.view_shard(sc.views, rank, world)
        f = torch.from_numpy(sc.features).to(dev).to(dtype).unsqueeze(1).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        p = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
        args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
        v1, c1, m1 = cn.aggregate_views(p, f, *args)                       # all views on this GPU
        scale = float(v1.abs().max())
        ex = D.ViewExchange(sc.views, channels, sc.height, sc.width, dtype, dev)
        for it in range(3):                                                # repeated calls reuse the symmetric buffers
            ex.local_features().copy_(f[lo:hi] * (1.0 if it == 2 else 0.5))   # earlier rounds leave stale rows behind
            for overlap in (False, True):
                blo, bdim, vol, cnt, valid = ex.aggregate(p, *args, overlap=overlap)
                if it < 2:
                    continue
                sl = tuple(slice(l, l + d) for l, d in zip(blo, bdim))
                good = torch.equal(cnt[0, 0], c1[0, 0][sl]) and torch.equal(valid[0, 0], m1[0, 0][sl])
                ref = v1[0][(slice(None),) + sl]
                # all views in view order, whatever the pipelining: the single-GPU bits
                good = good and torch.equal(vol[0].contiguous().view(torch.int32), ref.contiguous().view(torch.int32))
                ok, why = ok and good, why or ("" if good else f"C={channels} overlap={overlap}")
        ok = ok and ex.pulled_bytes() >= 0
    ret[f"ok{rank}"] = ok
    ret[f"why{rank}"] = why
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", WORLDS)
def test_voxel_sharded_feature_exchange(world):
    """ViewExchange: views in by rank, boxes of the volume out by rank; feature rows pulled over NVLink by the TMA
    puller.  The result is bit-identical to one GPU, with or without the part-by-part overlap."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29900 + (os.getpid() % 1000) + world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_exchange, args=(world, port, ret), nprocs=world, join=True)
        got = dict(ret)
    assert all(got.get(f"ok{r}") for r in range(world)), got
