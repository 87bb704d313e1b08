"""Multi-GPU partitioning of the aggregation path: one process per GPU (torch.distributed, NCCL over
NVLink / NVSwitch), two ways (the reference itself only has scene data-parallelism,
dist_train.sh:8):

  by scene   independent scenes, no data-path collective            -> scene_shard()
  by view    each rank lifts a contiguous range of views of ONE scene over the full grid; the partial
             sums and counts / weight totals are combined by ONE all-reduce of a packed fp32 buffer
             [sums | totals], then divided locally                  -> aggregate_views_sharded(),
                                                                       dense_rma_sharded(), rma_points_sharded()
             (chunks=K pipelines the collective against the kernel, x-range by x-range;
             collective="reduce_scatter" leaves every rank with 1/G of the voxels: half the wire bytes)
  by view in, by voxel out
             the views arrive sharded (each rank ran the 2D network on its own views) but every rank
             owns one BOX of the volume and lifts ALL views into it: no partial sums on the wire, the
             fp32 sums keep the reference's view order (bit-identical to one GPU), and what crosses
             NVLink are only the feature rows a box's voxels project to, pulled from the peers' symmetric
             memory by a TMA kernel that runs beside the gather kernel       -> ViewExchange

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


def default_splits(world_size):
    """(gx, gy, gz) with gx*gy*gz == world_size for box_shard: factors of two go to z, x, y in turn (boxes that are
    short in z and x see the smallest part of every image: fewest feature rows to exchange, profiles/r02_multi_gpu.md),
    what is left to x."""
    g = [1, 1, 1]                      # x, y, z
    order, i, w = (2, 0, 1), 0, int(world_size)
    while w % 2 == 0 and w > 1:
        g[order[i % 3]] *= 2
        w //= 2
        i += 1
    g[0] *= w
    return tuple(g)


def box_shard(voxel_dim, rank, world_size, splits=None):
    """The box of the volume owned by `rank`: (lo, dim), each a 3-tuple.  The grid is cut into gx x gy x gz boxes
    (default_splits), rank = (ix*gy + iy)*gz + iz; cuts at n*i//g, so the boxes tile the grid exactly."""
    gx, gy, gz = splits or default_splits(world_size)
    if gx * gy * gz != world_size:
        raise ValueError("splits must multiply to the world size")
    iz = rank % gz
    iy = (rank // gz) % gy
    ix = rank // (gz * gy)
    lo, dim = [], []
    for n, g, i in zip(voxel_dim, (gx, gy, gz), (ix, iy, iz)):
        a, b = int(n) * i // g, int(n) * (i + 1) // g
        lo.append(a)
        dim.append(b - a)
    if min(dim) <= 0:
        raise ValueError("more cuts than voxels along an axis")
    return tuple(lo), tuple(dim)


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
                            local_sums=None, finalize=None, batch=None, channels=None, device=None, chunks=1,
                            collective="all_reduce"):
    """View-sharded Stage A.  `projections` / `features` hold THIS RANK's views only (see view_shard).

    Returns (volume [B,C,nx,ny,nz] mean over all ranks' views, count [B,1,nx,ny,nz] fp32, valid bool),
    identical on every rank.  One all-reduce of B*nvox*(C+1) floats; the kernel writes sums and (float)
    counts directly into that buffer and a finalise-only launch divides in place afterwards.
    `chunks` > 1 or collective="reduce_scatter" select the pipelined forms (aggregate_views_pipelined)."""
    if chunks > 1 or collective != "all_reduce":
        return aggregate_views_pipelined(projections, features, voxel_dim, voxel_size, origin, stride, group=group,
                                         chunks=chunks, collective=collective)
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


_comm_streams = {}


def _comm_stream(device):
    st = _comm_streams.get(device)
    if st is None:
        st = _comm_streams[device] = torch.cuda.Stream(device=device)
    return st


def x_chunks(nx, parts):
    """`parts` contiguous x-ranges [(x0, x1), ...] covering [0, nx) (sizes differ by at most one, empty ones dropped)."""
    cuts = [int(nx) * i // parts for i in range(parts + 1)]
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


def aggregate_views_pipelined(projections, features, voxel_dim, voxel_size, origin, stride, group=None, chunks=4,
                              collective="all_reduce"):
    """View-sharded Stage A with the NCCL collective pipelined against the gather kernel (batch 1).

    The volume is cut into x-ranges; the kernel lifts this rank's views into one range at a time (a box launch:
    cnrma_aggregate_views_box) and the range's partial sums and float counts go out on a second stream while the next
    range is being computed; a finalise launch divides each range once its collective is done.
      collective="all_reduce"      every rank ends with the whole mean volume: returns (volume [1,C,nx,ny,nz],
                                   count fp32 [1,1,nx,ny,nz], valid bool), like aggregate_views_sharded
      collective="reduce_scatter"  every rank owns one x-slab (x_chunks(nx, world)), cut `chunks` times; each range is
                                   reduced to its owner only (ncclReduce) -- half the wire bytes of the all-reduce:
                                   returns (x0, x1, volume [1,C,x1-x0,ny,nz], count, valid) for this rank's slab
    Sums are regrouped by rank: 1e-5 like every view-sharded mode; indices, masks and counts exact."""
    rank, world = _world(group)
    nx, ny, nz = (int(v) for v in voxel_dim)
    f0 = features[0]
    if f0.shape[0] != 1:
        raise ValueError("the pipelined forms handle one scene at a time")
    channels, device = f0.shape[1], f0.device
    if collective not in ("all_reduce", "reduce_scatter"):
        raise ValueError("collective must be 'all_reduce' or 'reduce_scatter'")
    scatter = collective == "reduce_scatter"
    if scatter:
        slabs = x_chunks(nx, world)
        if len(slabs) != world:
            raise ValueError("fewer x-planes than ranks")
        ranges = [(a + c0, a + c1, o) for o, (a, b) in enumerate(slabs) for c0, c1 in x_chunks(b - a, max(1, chunks))]
        ox0, ox1 = slabs[rank]
    else:
        ranges = [(a, b, rank) for a, b in x_chunks(nx, max(1, chunks))]
        ox0, ox1 = 0, nx
    plane = ny * nz
    out_vol = torch.empty((1, ox1 - ox0, ny, nz, channels), dtype=torch.float32, device=device)
    out_cnt = torch.empty((1, 1, ox1 - ox0, ny, nz), dtype=torch.float32, device=device)
    main, comm = torch.cuda.current_stream(device), _comm_stream(device)
    mine = []
    for a, b, owner in ranges:
        if owner == rank:       # reduce in place: an x-range is a contiguous slice of both outputs
            vol, cnt = out_vol[:, a - ox0:b - ox0], out_cnt[:, :, a - ox0:b - ox0]
        else:
            vol = torch.empty((1, b - a, ny, nz, channels), dtype=torch.float32, device=device)
            cnt = torch.empty((1, 1, b - a, ny, nz), dtype=torch.float32, device=device)
        if len(projections) == 0:
            vol.zero_()
            cnt.zero_()
        else:
            F.aggregate_views(projections, features, voxel_dim, voxel_size, origin, stride, mean=False,
                              out=(vol.permute(0, 4, 1, 2, 3), cnt, None), accumulate=False, count_f32=True,
                              box=((a, 0, 0), (b - a, ny, nz)))
        ready = torch.cuda.Event()
        ready.record(main)
        comm.wait_event(ready)
        done = torch.cuda.Event()
        with torch.cuda.stream(comm):
            if world > 1:
                for t in (vol, cnt):
                    flat = t.reshape(-1)          # contiguous by construction: a view, not a copy
                    assert flat.data_ptr() == t.data_ptr()
                    if scatter:
                        dst = dist.get_global_rank(group, owner) if group is not None else owner
                        dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group)
                    else:
                        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            done.record(comm)
        vol.record_stream(comm)
        cnt.record_stream(comm)
        if owner == rank:
            mine.append((vol, cnt, done))
    for vol, cnt, done in mine:
        main.wait_event(done)
        F.finalize_views(vol.permute(0, 4, 1, 2, 3), cnt, count_f32=True)
    main.wait_stream(comm)
    volume = out_vol.permute(0, 4, 1, 2, 3)
    if scatter:
        return ox0, ox1, volume, out_cnt, out_cnt > 0
    return volume, out_cnt, out_cnt > 0


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

    def global_mean(result_block):
        """(sum of weights, M) of this rank's views, read from the march's result block on the device, all-reduced; no
        host round trip, so the fill is queued right behind the collective."""
        tot = torch.cat((result_block[8:16].view(torch.float64), result_block[0:8].view(torch.int64).to(torch.float64)))
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
        return (tot[0] / tot[1]).to(torch.float32).reshape(1)

    global_mean.on_device = True

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


def aggregate_views_p2p(projections, features, voxel_dim, voxel_size, origin, stride, group=None, channels=None,
                        device=None):
    """View-sharded Stage A with the exchange fused into the gather kernel (NVLink peer stores into symmetric memory,
    torch.distributed._symmetric_memory) instead of an all-reduce of full-size partial volumes.

    `projections` / `features` hold THIS RANK's views (view_shard).  The voxels are split into world_size contiguous
    ranges; every rank's kernel stores the partial sums of ALL voxels straight into their owners' buffers, a device
    barrier follows, and each owner adds what it received in rank order and divides.  Returns this rank's slab:
    (lo, hi, volume [hi - lo, C], count [hi - lo] int32, valid [hi - lo] bool), voxels in the flat order
    (x * ny + y) * nz + z.  All-gather the slabs if one rank needs the whole volume."""
    import torch.distributed._symmetric_memory as symm_mem
    rank, world = _world(group)
    if len(features) == 0 and (channels is None or device is None):
        raise ValueError("a rank without views must say `channels` and `device`")
    if len(features) > 0:
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
    if len(projections) == 0:                                 # a rank without views (V < world): its sections are zeros
        for o in range(world):
            peers[o][rank].zero_()
    else:
        route_views(projections, features, voxel_dim, voxel_size, origin, stride, [peers[o][rank] for o in range(world)])
    hdl.barrier(channel=1)                                    # all sources' stores have landed
    lo = min(rank * slab, nvox)
    hi = min(lo + slab, nvox)
    volume, count, valid = finalize_routed(buf, hi - lo, channels)
    return lo, hi, volume, count, valid


# ---------------------------------------------------------------------------------------------------------------
# Views in by rank, voxels out by rank: the feature rows a box needs are pulled from the peers over NVLink
# ---------------------------------------------------------------------------------------------------------------

class ViewExchange:
    """Voxel-sharded Stage A of ONE scene whose views arrive sharded over the ranks (batch 1).

    Every rank owns one box of the volume (box_shard) and lifts ALL views into it, in view order -- the sums are
    bit-identical to the single-GPU result and nothing but feature rows crosses NVLink:

      1. the producer (the 2D network) writes this rank's views into `local_features()`, a symmetric-memory buffer
         (torch.distributed._symmetric_memory) every peer can read;
      2. cnrma_mark_rows marks, per remote view, the pixel rows this rank's voxels project to (exact: the gather
         kernels' own projection arithmetic) -- a fifth to a third of them for the boxes and cameras at hand;
      3. cnrma_pull_rows copies exactly those rows out of the peers' buffers into a local staging copy (TMA bulk copies
         through shared memory, on a side stream, one launch per peer so that every peer serves one reader at a time);
      4. cnrma_aggregate_views_box gathers from the local views and the staged remote ones, all views in view order.
         With overlap=True the box is served part by part: the rows of the next part arrive while the current one is
         being gathered (the gather kernel leaves the puller its CTA slots).

    Returns this rank's box; all-gather the boxes if one rank needs the whole volume."""

    def __init__(self, views, channels, height, width, dtype, device, group=None, splits=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = _world(group)
        self.V, self.C, self.H, self.W = int(views), int(channels), int(height), int(width)
        self.dtype, self.device, self.splits = dtype, torch.device(device), splits
        self.shards = [view_shard(self.V, r, self.world) for r in range(self.world)]
        self.vmax = max(hi - lo for lo, hi in self.shards)
        shape = (self.vmax, self.H, self.W, self.C)
        self.local = symm_mem.empty(shape, dtype=dtype, device=self.device)
        self.hdl = symm_mem.rendezvous(self.local, self.group)
        self.peers = [self.hdl.get_buffer(q, shape, dtype) for q in range(self.world)]
        self.staging = torch.empty((self.V, self.H, self.W, self.C), dtype=dtype, device=self.device)
        self.words = (self.H * self.W + 31) // 32
        self._bm = torch.zeros((5, self.V, self.words), dtype=torch.int32, device=self.device)   # [0]: rows present
        self.bitmap = self._bm[0]
        self.side = torch.cuda.Stream(device=self.device, priority=-1)
        self.row_bytes = self.C * self.local.element_size()
        self.view_bytes = self.H * self.W * self.row_bytes
        self._descs = {}
        self._work = torch.zeros(16, dtype=torch.int32, device=self.device)      # one claim counter per part
        import ctypes as C
        ptrs = []
        for q, (qlo, qhi) in enumerate(self.shards):                             # view v lives in its owner's buffer
            ptrs += [self.peers[q][i].data_ptr() for i in range(qhi - qlo)]
        self._owner_ptrs = (C.c_void_p * self.V)(*ptrs)

    def local_features(self):
        """This rank's views as a [v_local, 1, C, H, W] tensor (channels-last rows in symmetric memory) to write into."""
        lo, hi = self.shards[self.rank]
        return self.local[: hi - lo].permute(0, 3, 1, 2).unsqueeze(1)

    def _descriptor(self, lo, hi):
        """cnrma_features over views [lo, hi): this rank's own buffer for its views, the staging copy for the others.
        The buffers are persistent, so the descriptors are built once."""
        import ctypes as C
        from . import _lib
        key = (lo, hi)
        d = self._descs.get(key)
        if d is None:
            mlo, mhi = self.shards[self.rank]
            ptrs = (C.c_void_p * (hi - lo))(*[(self.local[v - mlo] if mlo <= v < mhi else self.staging[v]).data_ptr()
                                              for v in range(lo, hi)])
            d = _lib.Features(hi - lo, self.C, self.H, self.W, F._DTYPES[self.dtype], 1, self.W * self.C, self.C,
                              C.cast(ptrs, C.POINTER(C.c_void_p)))
            d._keepalive = ptrs
            self._descs[key] = d
        return d

    def pulled_bytes(self):
        """Bytes the last aggregate() call read from the peers (host sync; for reports)."""
        mlo, mhi = self.shards[self.rank]
        bits = self.bitmap.cpu().numpy().view("uint8")
        import numpy as np
        per_view = np.unpackbits(bits, axis=1).sum(axis=1)
        per_view[mlo:mhi] = 0
        return int(per_view.sum()) * self.row_bytes

    def aggregate(self, projections, voxel_dim, voxel_size, origin, stride, mean=True, overlap=False, parts=4, pull_ctas=0,
                  pull_path=None, profile=None):
        """projections [V,1,3,4]: the cameras of ALL views (replicated; 48 bytes each).  The local views must already be
        in local_features() (written on the current stream).  Returns (lo, dim, volume [1,C,*dim], count int32
        [1,1,*dim], valid bool) for this rank's box -- bit-identical to the same voxels of a single-GPU call.

        overlap=False  mark, pull everything, then one gather launch over the box.
        overlap=True   the box is cut into `parts` x-ranges; the rows of part k+1 (those not already there) are pulled
                       while part k is being gathered, and the gather kernel leaves the puller its CTA slots.
        pull_path: "tma" (bulk copies through shared memory) or "lsu" (16-byte loads / stores through registers);
                   default: "lsu" beside the gather kernel (overlap), "tma" alone.
        profile: a dict that receives the device times (ms) of the phases of this call (host sync; for reports)."""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        dev = self.device
        P = F._projections_device(projections, dev)
        if P.shape[0] != self.V or P.shape[1] != 1:
            raise ValueError("projections must hold the cameras of all views of one scene: [V,1,3,4]")
        lo, dim = box_shard(voxel_dim, self.rank, self.world, self.splits)
        grid = _lib.make_grid(voxel_dim, voxel_size, F._origin3(origin))
        main = torch.cuda.current_stream(dev)
        lsu = (pull_path or ("lsu" if overlap else "tma")) == "lsu"
        if not pull_ctas and not overlap:
            pull_ctas = 2 * lib.cnrma_pull_default_ctas()    # nothing runs beside the puller: one CTA per SM
        pull_arg = (int(pull_ctas) & 0xFFFF) | (_lib.PULL_LSU if lsu else 0)
        order = [(self.rank + k) % self.world for k in range(1, self.world)]     # every peer serves one reader at a time
        order = [q for q in order if self.shards[q][1] > self.shards[q][0]]
        bx, by, bz = dim
        xparts = x_chunks(bx, max(1, min(int(parts), bx, 16))) if (overlap and order) else [(0, bx)]
        if self._bm.shape[0] < len(xparts) + 1:
            self._bm = torch.zeros((len(xparts) + 1, self.V, self.words), dtype=torch.int32, device=dev)
        done = self._bm[0]                                   # rows present in the staging copy
        self.bitmap = done
        mlo, mhi = self.shards[self.rank]
        remote = [(a, b) for a, b in ((0, mlo), (mhi, self.V)) if b > a]
        marks = []                                           # profiling: (label, event) on the side / main streams

        def stamp(stream, label):
            if profile is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                marks.append((label, ev))

        with torch.cuda.device(dev):
            stamp(main, "start")
            entered = torch.cuda.Event()
            entered.record(main)
            self.side.wait_event(entered)
            arrived = []
            with torch.cuda.stream(self.side):
                if order:
                    # which rows this box needs is a matter of cameras and geometry only: the marks run BEFORE the
                    # barrier, beside whatever the peers are still doing
                    self._bm[: len(xparts) + 1].zero_()
                    self._work.zero_()
                    box = _lib.make_box(lo, dim)              # the rows every part needs from the remote views
                    for a, b in remote:
                        _lib.check(lib.cnrma_mark_rows(C.byref(grid), C.byref(box), C.c_void_p(P[a].data_ptr()), 12,
                                                       b - a, float(stride), self.H, self.W,
                                                       C.c_void_p(self._bm[1, a].data_ptr()), len(xparts),
                                                       self.V * self.words, F._stream(dev)), "cnrma_mark_rows")
                    stamp(self.side, "marks")
            self.hdl.barrier(channel=0)                       # every rank's views are in its symmetric buffer
            start = torch.cuda.Event()
            start.record(main)
            stamp(main, "barrier")
            self.side.wait_event(start)
            with torch.cuda.stream(self.side):
                for k in range(len(xparts) if order else 0):
                    _lib.check(lib.cnrma_pull_rows(C.c_void_p(self._bm[k + 1].data_ptr()), C.c_void_p(done.data_ptr()),
                                                   self.V, self.H, self.W, self.row_bytes, self._owner_ptrs,
                                                   C.c_void_p(self.staging.data_ptr()), self.view_bytes, pull_arg,
                                                   C.c_void_p(self._work[k].data_ptr()), mhi % self.V, F._stream(dev)),
                               "cnrma_pull_rows")
                    ev = torch.cuda.Event()
                    ev.record(self.side)
                    arrived.append(ev)
                    stamp(self.side, f"pull{k}")
                if not order:
                    ev = torch.cuda.Event()
                    ev.record(self.side)
                    arrived.append(ev)
                self.hdl.barrier(channel=1)                   # every rank is done reading its peers' buffers
                released = torch.cuda.Event()
                released.record(self.side)
            buf = _lib.empty((1, bx, by, bz, self.C), dtype=torch.float32, device=dev)
            count = _lib.empty((1, 1, bx, by, bz), dtype=torch.int32, device=dev)
            valid = _lib.empty((1, 1, bx, by, bz), dtype=torch.bool, device=dev)
            desc = self._descriptor(0, self.V)
            flags = _lib.AGG_MEAN if mean else 0
            reserve = (int(pull_ctas) if pull_ctas else lib.cnrma_pull_default_ctas()) if len(xparts) > 1 else 0
            plane = by * bz
            for k, (x0, x1) in enumerate(xparts):
                main.wait_event(arrived[k])
                box = _lib.make_box((lo[0] + x0, lo[1], lo[2]), (x1 - x0, by, bz))
                _lib.check(lib.cnrma_aggregate_views_box(
                    C.byref(grid), C.byref(box), C.byref(desc), C.c_void_p(P.data_ptr()), 12, float(stride), flags,
                    C.c_void_p(buf.data_ptr() + x0 * plane * self.C * 4), self.C, 1,
                    C.c_void_p(count.data_ptr() + x0 * plane * 4), C.c_void_p(valid.data_ptr() + x0 * plane),
                    reserve if k + 1 < len(xparts) else 0, F._stream(dev)), "cnrma_aggregate_views_box")
                stamp(main, f"gather{k}")
            main.wait_event(released)                         # the next call may overwrite local_features()
            stamp(main, "end")
        if profile is not None:
            torch.cuda.synchronize(dev)
            t0 = marks[0][1]
            profile.update({label: t0.elapsed_time(ev) for label, ev in marks[1:]})
        return lo, dim, buf.permute(0, 4, 1, 2, 3), count, valid
