"""Times the backward of both lifts on cfg 2 (CUDA events, after warm-up): python profiles/microbench/backward_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cnrma_b200 as cn

sc = cn.synthetic.make_scene(sys.argv[1] if len(sys.argv) > 1 else "cfg2", seed=0, with_features=False)
dev = torch.device("cuda")
feats = cn.synthetic.device_features(sc, dev, channels_last=True).requires_grad_(True)
proj = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
tsdf = torch.from_numpy(sc.tsdf).to(dev)[None, None]
args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)


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


vol, cnt, _ = cn.aggregate_views(proj, feats, *args)
gv = torch.randn_like(vol)
pts = cn.rma_points(proj, feats, tsdf, *args, threshold=0.05)[0]
gp = torch.randn_like(pts)


def bwd_a():
    feats.grad = None
    vol.backward(gv, retain_graph=True)


def bwd_b():
    feats.grad = None
    pts.backward(gp, retain_graph=True)


ta, tb = timed(bwd_a), timed(bwd_b)
nb_a = vol.numel() * 4 + feats.numel() * 4 * 2          # read grad volume, zero + write grad features
nb_b = pts.numel() * 4 + feats.numel() * 4
print(f"stage A backward {ta:.3f} ms  ({nb_a / ta / 1e6:.0f} GB/s of compulsory bytes)")
print(f"stage B backward {tb:.3f} ms  ({nb_b / tb / 1e6:.0f} GB/s of compulsory bytes)")
