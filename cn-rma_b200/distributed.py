"""Multi-GPU partitioning of the aggregation path: one process per GPU (torch.distributed, NCCL over
NVLink / NVSwitch), two ways (the reference itself only has scene data-parallelism,
dist_train.sh:8):

  by scene   independent scenes, no data-path collective            -> scene_shard()
  by view    each rank lifts a contiguous range of views of ONE scene over the full grid; the partial
             sums and counts / weight totals are combined by ONE all-reduce of a packed fp32 buffer
             [sums | totals], then divided locally                  -> aggregate_views_sharded(),
                                                                       dense_rma_sharded(), rma_points_sharded()

Contiguous view ranges keep every rank's additions in view order, so per-rank partial sums are bit-exact;
the cross-rank sum order is NCCL's (tolerance 1e-5, indices / masks / counts still exact).

The compute hooks default to the CUDA library; the world_size-2 gloo tests inject oracle-backed hooks to
exercise the partitioning and packing logic on CPU.
"""
import torch
import torch.distributed as dist

from . import functional as F


def view_shard(num_views, rank, world_size):
    """Contiguous range [lo, hi) of views owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_views, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scene_shard(num_scenes, rank, world_size):
    """Scenes owned by `rank` under round-robin data parallelism (scene i -> GPU i mod G)."""
    return list(range(rank, num_scenes, world_size))


def _world(group):
    if not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _cuda_local_sums(projections, features, voxel_dim, voxel_size, origin, stride, packed):
    """Partial sums / counts of this rank's views, written by the kernel straight into the packed buffer."""
    vol_view, cnt_view = packed
    if len(projections) == 0:
        vol_view.zero_()
        cnt_view.zero_()
        return
    F.aggregate_views(projections, features, voxel_dim, voxel_size, origin, stride, mean=False,
                      out=(vol_view, cnt_view, None), accumulate=False, count_f32=True)


def _cuda_finalize(vol_view, cnt_view):
    return F.finalize_views(vol_view, cnt_view, count_f32=True)


def aggregate_views_sharded(projections, features, voxel_dim, voxel_size, origin, stride, group=None,
                            local_sums=None, finalize=None, batch=None, channels=None, device=None):
    """View-sharded Stage A.  `projections` / `features` hold THIS RANK's views only (see view_shard).

    Returns (volume [B,C,nx,ny,nz] mean over all ranks' views, count [B,1,nx,ny,nz] fp32, valid bool),
    identical on every rank.  One all-reduce of B*nvox*(C+1) floats; the kernel writes sums and (float)
    counts directly into that buffer and a finalise-only launch divides in place afterwards."""
    nx, ny, nz = (int(v) for v in voxel_dim)
    if batch is None:
        f0 = features[0]
        batch, channels, device = f0.shape[0], f0.shape[1], f0.device
    nvox = nx * ny * nz
    buf = torch.empty(batch * nvox * (channels + 1), dtype=torch.float32, device=device)
    vol_view = buf[: batch * nvox * channels].view(batch, nx, ny, nz, channels).permute(0, 4, 1, 2, 3)
    cnt_view = buf[batch * nvox * channels:].view(batch, 1, nx, ny, nz)
    (local_sums or _cuda_local_sums)(projections, features, voxel_dim, voxel_size, origin, stride,
                                     (vol_view, cnt_view))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)      # counts as fp32 are exact below 2**24
    volume = (finalize or _cuda_finalize)(vol_view, cnt_view)
    return volume, cnt_view, cnt_view > 0


def dense_rma_sharded(projections, features, tsdf, voxel_dim, voxel_size, origin, stride, grids=300, mode="neus",
                      threshold=None, depth_points=None, group=None, local_scatter=None, batch=None, channels=None,
                      device=None):
    """View-sharded dense RMA: partial (wsum, wtot) per rank, ONE all-reduce of the packed buffer.
    Returns un-normalised (wsum [B,C,nx,ny,nz], wtot [B,1,nx,ny,nz]) identical on every rank."""
    nx, ny, nz = (int(v) for v in voxel_dim)
    if batch is None:
        f0 = features[0]
        batch, channels, device = f0.shape[0], f0.shape[1], f0.device
    nvox = nx * ny * nz
    buf = torch.zeros(batch * nvox * (channels + 1), dtype=torch.float32, device=device)
    wsum = buf[: batch * nvox * channels].view(batch, nx, ny, nz, channels).permute(0, 4, 1, 2, 3)
    wtot = buf[batch * nvox * channels:].view(batch, 1, nx, ny, nz)
    if local_scatter is not None:
        local_scatter(projections, features, tsdf, (wsum, wtot))
    elif len(projections) > 0:
        F.dense_rma(projections, features, tsdf, voxel_dim, voxel_size, origin, stride, grids=grids, mode=mode,
                    threshold=threshold, depth_points=depth_points, out=(wsum, wtot))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return wsum, wtot


def rma_points_sharded(projections, features, tsdf, voxel_dim, voxel_size, origin, stride, grids=300, mode="neus",
                       threshold=None, depth_points=None, group=None):
    """View-sharded point form: each rank emits the rows of its own views, normalised by the GLOBAL mean
    weight (all-reduce of two scalars: sum(w), M) which the fill kernel takes as its divisor.  Rank order
    == view order, so concatenating the ranks' outputs reproduces the single-GPU row order."""
    _rank, world = _world(group)

    def global_mean(weight_sum, rows, device):
        tot = torch.tensor([weight_sum, float(rows)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
        return (tot[0] / tot[1]).to(torch.float32).reshape(1)

    return F.rma_points(projections, features, tsdf, voxel_dim, voxel_size, origin, stride, grids=grids, mode=mode,
                        threshold=threshold, depth_points=depth_points, normalize=True, mean_hook=global_mean)


# ---------------------------------------------------------------------------------------------------------------
# View-sharded Stage A over peer memory: the exchange is done by the gather kernel's own stores
# ---------------------------------------------------------------------------------------------------------------

def _routed_slab(nvox, owners):
    return -(-nvox // owners)


def route_views(projections, features, voxel_dim, voxel_size, origin, stride, owner_rows):
    """This rank's views -> un-normalised sums and counts, stored by the kernel into `owner_rows[o]`, this source's
    [slab, C + 4] section of owner o's buffer (a peer-mapped tensor when o is another GPU).  Batch 1."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    fs = F._FeatureStack(F._as_view_list(features), need_vector_layout=True)
    if fs.B != 1:
        raise ValueError("route_views handles one scene at a time")
    P = F._projections_device(projections, fs.device)
    nx, ny, nz = (int(v) for v in voxel_dim)
    slab, row_floats = owner_rows[0].shape[0], owner_rows[0].shape[1]
    grid = _lib.make_grid(voxel_dim, voxel_size, F._origin3(origin))
    ptrs = (C.c_void_p * len(owner_rows))(*[t.data_ptr() for t in owner_rows])
    desc = fs.descriptor(0)
    with torch.cuda.device(fs.device):
        _lib.check(lib.cnrma_aggregate_views_routed(C.byref(grid), C.byref(desc), C.c_void_p(P[0, 0].data_ptr()), 12,
                                                    float(stride), len(owner_rows), slab, row_floats, ptrs,
                                                    F._stream(fs.device)), "cnrma_aggregate_views_routed")


def finalize_routed(recv, rows, channels, mean=True):
    """recv [n_src, slab, C + 4] (this owner's buffer, every source's section filled) -> (volume [rows, C] f32,
    count [rows] int32, valid [rows] bool): sums and counts added in source order, then the mean of rm.py:251."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    n_src, slab, row_floats = recv.shape
    volume = _lib.empty((rows, channels), dtype=torch.float32, device=recv.device)
    count = _lib.empty((rows,), dtype=torch.int32, device=recv.device)
    valid = _lib.empty((rows,), dtype=torch.bool, device=recv.device)
    with torch.cuda.device(recv.device):
        _lib.check(lib.cnrma_finalize_routed(C.c_void_p(recv.data_ptr()), n_src, slab, row_floats, rows, channels,
                                             1 if mean else 0, C.c_void_p(volume.data_ptr()), C.c_void_p(count.data_ptr()),
                                             C.c_void_p(valid.data_ptr()), F._stream(recv.device)), "cnrma_finalize_routed")
    return volume, count, valid


_p2p_buffers = {}


def aggregate_views_p2p(projections, features, voxel_dim, voxel_size, origin, stride, group=None):
    """View-sharded Stage A with the exchange fused into the gather kernel (NVLink peer stores into symmetric memory,
    torch.distributed._symmetric_memory) instead of an all-reduce of full-size partial volumes.

    `projections` / `features` hold THIS RANK's views (view_shard).  The voxels are split into world_size contiguous
    ranges; every rank's kernel stores the partial sums of ALL voxels straight into their owners' buffers, a device
    barrier follows, and each owner adds what it received in rank order and divides.  Returns this rank's slab:
    (lo, hi, volume [hi - lo, C], count [hi - lo] int32, valid [hi - lo] bool), voxels in the flat order
    (x * ny + y) * nz + z.  All-gather the slabs if one rank needs the whole volume."""
    import torch.distributed._symmetric_memory as symm_mem
    rank, world = _world(group)
    f0 = features[0]
    channels, device = f0.shape[1], f0.device
    nx, ny, nz = (int(v) for v in voxel_dim)
    nvox = nx * ny * nz
    slab, row_floats = _routed_slab(nvox, world), channels + 4
    key = (world, slab, row_floats, device)
    if key not in _p2p_buffers:
        buf = symm_mem.empty((world, slab, row_floats), dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
        peers = [hdl.get_buffer(o, (world, slab, row_floats), torch.float32) for o in range(world)]
        _p2p_buffers[key] = (buf, hdl, peers)
    buf, hdl, peers = _p2p_buffers[key]
    hdl.barrier(channel=0)                                    # every owner is done reading the previous call's data
    route_views(projections, features, voxel_dim, voxel_size, origin, stride, [peers[o][rank] for o in range(world)])
    hdl.barrier(channel=1)                                    # all sources' stores have landed
    lo = min(rank * slab, nvox)
    hi = min(lo + slab, nvox)
    volume, count, valid = finalize_routed(buf, hi - lo, channels)
    return lo, hi, volume, count, valid
