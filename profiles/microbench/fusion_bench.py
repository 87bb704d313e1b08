"""GT TSDF fusion throughput: 256x256x96 @ 0.04 m volume, 640x480 depth + colour frames (ScanNet-shaped)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cnrma_b200 as cn
from cnrma_b200 import synthetic

dim, vs, F, H, W = (256, 256, 96), 0.04, 100, 480, 640
rng = np.random.default_rng(0)
extent = tuple(d * vs for d in dim)
P, k, poses = synthetic.ring_cameras(F, H, W, 1, extent, rng, return_poses=True)
depth = torch.from_numpy(synthetic.room_depth_maps(k, poses, H, W, extent, rng)).cuda()
color = torch.rand((F, 3, H, W), device="cuda") * 255
P = torch.from_numpy(P).cuda()
fus = cn.TSDFFusion(dim, vs, (0, 0, 0), device="cuda", color=True)
for _ in range(2):
    fus.reset()
    fus.integrate_frames(P, depth, color)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fus.reset()
a.record()
fus.integrate_frames(P, depth, color)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
nvox = dim[0] * dim[1] * dim[2]
print(f"{F} frames into {nvox / 1e6:.1f} M voxels: {ms:.2f} ms = {ms / F * 1e3:.0f} us/frame, "
      f"{F * nvox / ms / 1e6:.1f} G voxel*frames/s; observed voxels {int((fus.weight_vol > 0).sum())}")
