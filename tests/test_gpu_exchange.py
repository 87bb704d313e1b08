"""GPU (one device): the building blocks of the voxel-sharded multi-GPU Stage A -- box launches, the needed-row
bitmap and the row puller -- each against the full-grid single-launch result, which test_gpu_parity.py pins to the
oracle and the reference's golden vectors.  The multi-rank protocol itself runs in tests/test_gpu_multi.py."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cn():
    import cnrma_b200
    cnrma_b200.load()
    return cnrma_b200


def _scene(cn, channels, seed=31, views=7, name="small"):
    sc = cn.synthetic.make_scene(name, seed=seed, channels=channels, views=views)
    f = torch.from_numpy(sc.features).cuda().unsqueeze(1).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    p = torch.from_numpy(sc.projections).cuda().unsqueeze(1)
    return sc, f, p


def _boxes(dim):
    nx, ny, nz = dim
    return [((0, 0, 0), dim), ((0, 0, nz // 2), (nx, ny, nz - nz // 2)), ((nx // 3, 1, 0), (nx - nx // 3, ny - 2, nz // 2)),
            ((nx - 1, ny - 1, nz - 1), (1, 1, 1)), ((2, 3, 1), (5, 4, 3))]


@pytest.mark.parametrize("channels", [16, 256])      # list kernel / TMA kernel
def test_box_launch_reproduces_the_full_grid_bits(cn, channels):
    sc, f, p = _scene(cn, channels)
    args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    vol, cnt, valid = cn.aggregate_views(p, f, *args, mean=True)
    for lo, dim in _boxes(sc.voxel_dim):
        sl = tuple(slice(l, l + d) for l, d in zip(lo, dim))
        bv, bc, bm = cn.aggregate_views(p, f, *args, mean=True, box=(lo, dim))
        assert tuple(bv.shape) == (1, channels) + tuple(dim)
        assert torch.equal(bc[0, 0], cnt[0, 0][sl]) and torch.equal(bm[0, 0], valid[0, 0][sl])
        assert torch.equal(bv[0].view(torch.int32), vol[0][(slice(None),) + sl].contiguous().view(torch.int32))
        # view chunks accumulated into the box in view order: still the same bits
        out = (torch.empty_like(bv), torch.empty_like(bc), torch.empty_like(bm))
        cn.aggregate_views(p[:3], f[:3], *args, mean=False, out=out, accumulate=False, box=(lo, dim))
        cn.aggregate_views(p[3:], f[3:], *args, mean=True, out=out, accumulate=True, box=(lo, dim), reserve_ctas=40)
        assert torch.equal(out[1], bc) and torch.equal(out[0].view(torch.int32), bv.view(torch.int32))


def _mark(cn, sc, p, lo, dim, H, W):
    from cnrma_b200 import _lib
    lib = cn.load()
    words = (H * W + 31) // 32
    bitmap = torch.zeros((sc.views, words), dtype=torch.int32, device="cuda")
    grid = _lib.make_grid(sc.voxel_dim, sc.voxel_size, sc.origin)
    box = _lib.make_box(lo, dim)
    _lib.check(lib.cnrma_mark_rows(C.byref(grid), C.byref(box), C.c_void_p(p.data_ptr()), 12, sc.views, float(sc.stride),
                                   H, W, C.c_void_p(bitmap.data_ptr()), 1, 0, None), "cnrma_mark_rows")
    return bitmap


def test_mark_rows_in_parts_equals_separate_marks(cn):
    from cnrma_b200 import _lib, distributed as D
    lib = cn.load()
    sc, f, p = _scene(cn, 16)
    H, W = sc.height, sc.width
    words = (H * W + 31) // 32
    grid = _lib.make_grid(sc.voxel_dim, sc.voxel_size, sc.origin)
    lo, dim = (3, 1, 2), (11, 17, 5)
    for parts in (2, 3, 11):
        bm = torch.zeros((parts, sc.views, words), dtype=torch.int32, device="cuda")
        box = _lib.make_box(lo, dim)
        _lib.check(lib.cnrma_mark_rows(C.byref(grid), C.byref(box), C.c_void_p(p.data_ptr()), 12, sc.views, float(sc.stride),
                                       H, W, C.c_void_p(bm.data_ptr()), parts, sc.views * words, None), "cnrma_mark_rows")
        for k, (x0, x1) in enumerate(D.x_chunks(dim[0], parts)):
            one = _mark(cn, sc, p, (lo[0] + x0, lo[1], lo[2]), (x1 - x0, dim[1], dim[2]), H, W)
            assert torch.equal(bm[k], one), (parts, k)


def _ptrs(t):
    """HOST array of the device pointers of t[0], t[1], ... (cnrma_pull_rows' src_view_ptrs_host)."""
    return (C.c_void_p * t.shape[0])(*[t[v].data_ptr() for v in range(t.shape[0])])


def _bits(bitmap, pixels):
    b = np.unpackbits(bitmap.cpu().numpy().view(np.uint8), axis=1, bitorder="little")
    return b[:, :pixels].astype(bool)


def test_mark_rows_is_exactly_the_set_of_gathered_rows(cn):
    sc, f, p = _scene(cn, 16)
    H, W = sc.height, sc.width
    px, py, valid = cn.project_views(p, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, H, W)
    px, py, valid = (t[:, 0].cpu().numpy().reshape((sc.views,) + tuple(sc.voxel_dim)) for t in (px, py, valid))
    for lo, dim in _boxes(sc.voxel_dim):
        sl = (slice(None),) + tuple(slice(l, l + d) for l, d in zip(lo, dim))
        want = np.zeros((sc.views, H * W), bool)
        for v in range(sc.views):
            m = valid[sl][v]
            want[v, (py[sl][v][m] * W + px[sl][v][m])] = True
        got = _bits(_mark(cn, sc, p.contiguous(), lo, dim, H, W), H * W)
        assert np.array_equal(got, want), (lo, dim)


@pytest.mark.parametrize("channels,dtype", [(8, torch.float32), (32, torch.float32), (256, torch.float32),
                                            (128, torch.bfloat16), (2048, torch.float32)])
def test_pull_rows_copies_exactly_the_marked_rows(cn, channels, dtype):
    from cnrma_b200 import _lib
    lib = cn.load()
    V, H, W = 3, 9, 23                                  # 207 pixels: a partial last bitmap word
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randn((V, H, W, channels), device="cuda", generator=g).to(dtype)
    words = (H * W + 31) // 32
    for density in (0.0, 0.05, 0.5, 1.0):
        mask = torch.rand((V, H * W), device="cuda", generator=g) < density
        bits = np.zeros((V, words * 32), np.uint8)
        bits[:, :H * W] = mask.cpu().numpy()
        bits[:, H * W:] = 1                             # stray bits beyond the image must be ignored
        bitmap = torch.from_numpy(np.packbits(bits, axis=1, bitorder="little").view(np.int32).copy()).cuda()
        dst = torch.full_like(src, -7.0)
        row_bytes = channels * src.element_size()
        for ctas in (0, 1, 3, 0x10000, 0x10002):          # CNRMA_PULL_LSU: the load/store path
            dst.fill_(-7.0)
            work = torch.zeros(1, dtype=torch.int32, device="cuda")
            _lib.check(lib.cnrma_pull_rows(C.c_void_p(bitmap.data_ptr()), None, V, H, W, row_bytes, _ptrs(src),
                                           C.c_void_p(dst.data_ptr()), H * W * row_bytes, ctas,
                                           C.c_void_p(work.data_ptr()) if ctas != 3 else None, (ctas & 0xFFFF) % V, None), "cnrma_pull_rows")
            torch.cuda.synchronize()
            want = torch.where(mask.view(V, H, W, 1), src, torch.full_like(src, -7.0))
            assert torch.equal(dst, want), (channels, density, ctas)


def test_pull_rows_skips_and_records_rows_already_present(cn):
    from cnrma_b200 import _lib
    lib = cn.load()
    V, H, W, channels = 2, 8, 16, 64
    g = torch.Generator(device="cuda").manual_seed(6)
    src = torch.randn((V, H, W, channels), device="cuda", generator=g)
    a = torch.rand((V, H * W), device="cuda", generator=g) < 0.4
    b = torch.rand((V, H * W), device="cuda", generator=g) < 0.4
    pack = lambda m: torch.from_numpy(np.packbits(m.cpu().numpy().astype(np.uint8), axis=1, bitorder="little").view(np.int32).copy()).cuda()
    done = torch.zeros((V, H * W // 32), dtype=torch.int32, device="cuda")
    dst = torch.full_like(src, -7.0)
    rb = channels * 4
    call = lambda bm: _lib.check(lib.cnrma_pull_rows(C.c_void_p(bm.data_ptr()), C.c_void_p(done.data_ptr()), V, H, W, rb,
                                                     _ptrs(src), C.c_void_p(dst.data_ptr()), H * W * rb, 0, None, 1, None),
                                 "cnrma_pull_rows")
    call(pack(a))
    assert torch.equal(done, pack(a))
    dst[a.view(V, H, W)] = -3.0                         # if the second call pulled these again they would be restored
    call(pack(b))
    torch.cuda.synchronize()
    assert torch.equal(done, pack(a | b))
    want = torch.where((b & ~a).view(V, H, W, 1), src, torch.full_like(src, -7.0))
    want[a.view(V, H, W)] = -3.0
    assert torch.equal(dst, want)


@pytest.mark.parametrize("channels,world", [(16, 2), (256, 4), (256, 8), (32, 3)])
def test_exchange_on_one_gpu_with_simulated_ranks(cn, channels, world):
    """Every simulated rank marks the rows of its box, pulls them out of the owners' buffers into a NaN-poisoned
    staging copy and lifts all views from there: bit-identical to the single launch, so no needed row is missed."""
    from cnrma_b200 import _lib, distributed as D
    lib = cn.load()
    sc, f, p = _scene(cn, channels, views=9)
    H, W = sc.height, sc.width
    args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    vol, cnt, valid = cn.aggregate_views(p, f, *args, mean=True)
    rows = f[:, 0].permute(0, 2, 3, 1).contiguous()            # [V,H,W,C]: what the ranks' symmetric buffers hold
    row_bytes = channels * 4
    for rank in range(world):
        lo, dim = D.box_shard(sc.voxel_dim, rank, world)
        mlo, mhi = D.view_shard(sc.views, rank, world)
        bitmap = _mark(cn, sc, p.contiguous(), lo, dim, H, W)
        staging = torch.full_like(rows, float("nan"))
        for q in range(world):
            qlo, qhi = D.view_shard(sc.views, q, world)
            if q == rank or qhi == qlo:
                continue
            _lib.check(lib.cnrma_pull_rows(C.c_void_p(bitmap[qlo].data_ptr()), None, qhi - qlo, H, W, row_bytes,
                                           _ptrs(rows[qlo:qhi]), C.c_void_p(staging[qlo].data_ptr()), H * W * row_bytes, 0,
                                           None, 0, None), "cnrma_pull_rows")
        staging[mlo:mhi] = rows[mlo:mhi]
        fv = [staging[v].permute(2, 0, 1).unsqueeze(0) for v in range(sc.views)]
        bv, bc, bm = cn.aggregate_views(p, fv, *args, mean=True, box=(lo, dim))
        sl = tuple(slice(l, l + d) for l, d in zip(lo, dim))
        assert torch.equal(bc[0, 0], cnt[0, 0][sl])
        assert torch.equal(bv[0].view(torch.int32), vol[0][(slice(None),) + sl].contiguous().view(torch.int32)), rank


def test_box_and_exchange_argument_errors(cn):
    from cnrma_b200 import _lib
    lib = cn.load()
    sc, f, p = _scene(cn, 16)
    with pytest.raises(ValueError):
        cn.aggregate_views(p, f, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride, box=((0, 0, 0), (sc.voxel_dim[0] + 1, 1, 1)))
    bitmap = torch.zeros(64, dtype=torch.int32, device="cuda")
    buf = torch.zeros(4096, dtype=torch.float32, device="cuda")
    one = (C.c_void_p * 1)(buf.data_ptr())
    assert lib.cnrma_pull_rows(C.c_void_p(bitmap.data_ptr()), None, 1, 4, 4, 24, one, C.c_void_p(buf.data_ptr()), 384, 0,
                               None, 0, None) == -2                                     # row_bytes not a multiple of 16
    assert lib.cnrma_pull_rows(None, None, 1, 4, 4, 32, one, C.c_void_p(buf.data_ptr()), 512, 0, None, 0, None) == -1
    grid = _lib.make_grid(sc.voxel_dim, sc.voxel_size, sc.origin)
    bad = _lib.make_box((0, 0, 0), (sc.voxel_dim[0], sc.voxel_dim[1], sc.voxel_dim[2] + 1))
    assert lib.cnrma_mark_rows(C.byref(grid), C.byref(bad), C.c_void_p(p.data_ptr()), 12, 1, 4.0, 4, 4,
                               C.c_void_p(bitmap.data_ptr()), 1, 0, None) == -1
