"""One GPU: device times of the building blocks of the voxel-sharded Stage A (cnrma_mark_rows, cnrma_pull_rows with a
same-GPU source, the box gather) on cfg 4 geometry.  python profiles/microbench/exchange_parts_bench.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cnrma_b200 as cn
from cnrma_b200 import _lib, distributed as D

lib = cn.load()
dev = torch.device("cuda")
sc = cn.synthetic.make_scene("cfg4", seed=0, with_features=False)
full = cn.synthetic.device_features(sc, dev, channels_last=True)
rows = full[:, 0].permute(0, 2, 3, 1)                      # [V,H,W,C] contiguous
assert rows.is_contiguous()
P = torch.from_numpy(sc.projections).to(dev).unsqueeze(1).contiguous()
H, W, V = sc.height, sc.width, sc.views
words = (H * W + 31) // 32
grid = _lib.make_grid(sc.voxel_dim, sc.voxel_size, sc.origin)
row_bytes = sc.channels * 4


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for world in (2, 4, 8):
    lo, dim = D.box_shard(sc.voxel_dim, 0, world)
    box = _lib.make_box(lo, dim)
    bm = torch.zeros((V, words), dtype=torch.int32, device=dev)
    nv = V - V // world
    mark = lambda: _lib.check(lib.cnrma_mark_rows(C.byref(grid), C.byref(box), C.c_void_p(P[V // world].data_ptr()), 12, nv,
                                                  float(sc.stride), H, W, C.c_void_p(bm[V // world].data_ptr()), 1, 0, None), "mark")
    t_mark = timed(mark)
    bits = int(sum(bin(int(x) & 0xFFFFFFFF).count("1") for x in bm.cpu().numpy().ravel()))
    staging = torch.empty_like(rows)
    for ctas in (0, 37, 148, 296):
        ptrs = (C.c_void_p * V)(*[rows[v].data_ptr() for v in range(V)])
        pull = lambda: _lib.check(lib.cnrma_pull_rows(C.c_void_p(bm.data_ptr()), None, V, H, W, row_bytes, ptrs,
                                                      C.c_void_p(staging.data_ptr()), H * W * row_bytes, ctas, None, 0, None),
                                  "pull")
        t_pull = timed(pull)
        print(f"world {world}: box {dim}, mark {nv} views {t_mark*1e3:.1f} us; rows {bits} = {bits*row_bytes/1e6:.0f} MB; "
              f"local pull ctas={ctas}: {t_pull*1e3:.1f} us = {bits*row_bytes/t_pull/1e6:.0f} GB/s")
    args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    for reserve in (0, 74):
        t = timed(lambda: cn.aggregate_views(P, full, *args, box=(lo, dim), reserve_ctas=reserve))
        print(f"   box gather (all views local) reserve={reserve}: {t:.3f} ms")
