"""LRU model of the Stage A gather stream for different voxel sweep orders (CPU, numpy): python profiles/microbench/sweep_lru_sim.py cfg2"""
import numpy as np, sys, time
from collections import OrderedDict
sys.path.insert(0, "/root/repo")
import cnrma_b200.synthetic as S
sc = S.make_scene(sys.argv[1] if len(sys.argv) > 1 else "cfg2", seed=0, with_features=False)
P = sc.projections.astype(np.float32).copy(); P[:, :2] /= np.float32(sc.stride)
nx, ny, nz = sc.voxel_dim; vs = np.float32(sc.voxel_size); W, H, V = sc.width, sc.height, sc.views
gx, gy, gz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
pts = np.stack([gx.ravel() * vs, gy.ravel() * vs, gz.ravel() * vs, np.ones(gx.size, np.float32)], 0).astype(np.float32)
rows = []   # per voxel: list of row ids
ids = np.full((V, gx.size), -1, np.int64)
for v in range(V):
    cam = P[v] @ pts
    px = np.rint(cam[0] / cam[2]); py = np.rint(cam[1] / cam[2])
    ok = (cam[2] > 0) & (px >= 0) & (px < W) & (py >= 0) & (py < H)
    ids[v, ok] = (v * H + py[ok].astype(np.int64)) * W + px[ok].astype(np.int64)
print("gathers", int((ids >= 0).sum()), "distinct", len(np.unique(ids[ids >= 0])))
vox_index = (gx * ny + gy) * nz + gz   # flat id
def order_slab(T, inner="xy"):
    # slabs of T z-slices; within slab sweep x (slow), y, z (fast)
    o = []
    for z0 in range(0, nz, T):
        zz = np.arange(z0, min(nz, z0 + T))
        X, Y, Z = np.meshgrid(np.arange(nx), np.arange(ny), zz, indexing="ij")
        o.append(((X * ny + Y) * nz + Z).ravel())
    return np.concatenate(o)
def simulate(order, cap_rows, interleave=4736):
    # the kernel's warps take voxels it = warp + k * warps_total; all warps advance together, so the global gather
    # stream is approximately in `order` sequence
    cache = OrderedDict(); miss = 0; tot = 0
    sub = ids[:, order]
    for j in range(sub.shape[1]):
        col = sub[:, j]
        for r in col[col >= 0]:
            tot += 1
            if r in cache:
                cache.move_to_end(r)
            else:
                miss += 1
                cache[r] = None
                if len(cache) > cap_rows: cache.popitem(last=False)
    return miss, tot
for cap_mb in (48, 80):
    cap = cap_mb * 1024 * 1024 // (sc.channels * 4)
    for T in (1, 2, 4, 8, 32):
        t0 = time.time()
        miss, tot = simulate(order_slab(T), cap)
        print(f"L2 {cap_mb} MB  slab T={T:2d}: misses {miss} of {tot} = {miss * sc.channels * 4 / 1e9:.3f} GB read  ({time.time() - t0:.0f}s)")
print("---- variants at 64 MB")
cap = 64 * 1024 * 1024 // (sc.channels * 4)
def order_slab2(T, mode):
    o = []
    for z0 in range(0, nz, T):
        zz = np.arange(z0, min(nz, z0 + T))
        if mode == "x_y_z":
            X, Y, Z = np.meshgrid(np.arange(nx), np.arange(ny), zz, indexing="ij")
        elif mode == "x_z_y":
            X, Z, Y = np.meshgrid(np.arange(nx), zz, np.arange(ny), indexing="ij")
        elif mode == "tile8":   # 8x8 (x,y) tiles, all z of the slab inside a tile
            bx, by = np.meshgrid(np.arange(0, nx, 8), np.arange(0, ny, 8), indexing="ij")
            parts = []
            for tx, ty in zip(bx.ravel(), by.ravel()):
                Xs, Ys, Zs = np.meshgrid(np.arange(tx, min(nx, tx + 8)), np.arange(ty, min(ny, ty + 8)), zz, indexing="ij")
                parts.append(((Xs * ny + Ys) * nz + Zs).ravel())
            o.append(np.concatenate(parts)); continue
        o.append(((X * ny + Y) * nz + Z).ravel())
    return np.concatenate(o)
for T in (1, 4, 6, 8, 12, 16):
    for mode in ("x_y_z", "x_z_y", "tile8"):
        miss, tot = simulate(order_slab2(T, mode), cap)
        print(f"T={T:2d} {mode}: {miss * sc.channels * 4 / 1e9:.3f} GB read")
