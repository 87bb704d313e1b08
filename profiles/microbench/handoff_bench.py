"""cfg 2: full point cloud + switch_pointcloud (two steps) vs the fused selected fill, max_points = 500000 (rm.py config
:203).  The mask is drawn once outside the timed loops (np.random.choice over 5.76 M rows takes ~0.1-0.3 s on the host --
that cost is the reference's and is the same for both variants)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cnrma_b200 as cn

sc = cn.synthetic.make_scene("cfg2", seed=0, with_features=False)
dev = torch.device("cuda")
feats = cn.synthetic.device_features(sc, dev, channels_last=True)
proj = torch.from_numpy(sc.projections).to(dev).unsqueeze(1)
tsdf = torch.from_numpy(sc.tsdf).to(dev)[None, None]
args = (sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
off = [np.array([0.5, -0.25, 1.0], np.float32)]
full = cn.rma_points(proj, feats, tsdf, *args, threshold=0.05)[0]
mask = torch.from_numpy(cn.sample_points(full.shape[0], 500000, np.random.RandomState(0))).to(dev)


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


def two_steps():
    pts = cn.rma_points(proj, feats, tsdf, *args, threshold=0.05)
    return cn.switch_pointcloud(pts, off, masks=[mask])


def fused():
    return cn.rma_points_selected(proj, feats, tsdf, *args, offsets=off, masks=[mask], threshold=0.05)


a, b = two_steps(), fused()
assert torch.equal(a[0][0], b[0][0]) and torch.equal(a[1][0], b[1][0])
print(f"rows {full.shape[0]} -> {int(mask.sum())}")
print(f"rma_points + switch_pointcloud : {timed(two_steps):.3f} ms")
print(f"fused selected fill            : {timed(fused):.3f} ms")
print(f"switch_pointcloud alone        : {timed(lambda: cn.switch_pointcloud([full], off, masks=[mask])):.3f} ms")

import time
t0 = time.perf_counter()
cn.sample_points(full.shape[0], 500000, np.random.RandomState(1))
host_ms = (time.perf_counter() - t0) * 1e3
dev_ms = timed(lambda: cn.sample_points_device(full.shape[0], 500000, 1, dev))
print(f"mask draw: numpy on the host {host_ms:.1f} ms, on the device {dev_ms:.3f} ms")
