"""Seeded synthetic ScanNet/ARKit-shaped scenes (inputs only, no reference code).

Conventions follow the reference's data pipeline so that the inputs have the
shapes and statistics the aggregation path sees in production:

* `projection = K @ inv(E)[:3]` with `E` a camera-to-world pose, exactly as
  `projects/mvsdetection/datasets/pipelines/atlas_transforms.py:98-110`
  builds it; intrinsics are for the full-resolution image (4H x 4W) and the
  aggregation path divides rows 0-1 by `backbone2d_stride`
  (`ray_marching.py:238-239`, `:275-276`).
* TSDF sign convention of the Atlas fusion code
  (`data_prepare/scannet/tsdf.py:427-428`): negative in free space in front of
  a surface, positive behind it, +1 where unobserved; range [-1.05, 1.05]
  (`atlas_head.py:41`).
* `origin = [0, 0, 0]` (shipped configs, `ray_marching_scannet.py:30`).

Everything is generated with numpy on the host from an integer seed so the CPU
oracle and the CUDA path consume bit-identical inputs; `device_features`
offers an on-device generator for bench-sized feature stacks.
"""
from dataclasses import dataclass, field
import math

import numpy as np

# BASELINE.json `configs`, in order (SURVEY.md section 8d fills in the unstated sizes).
CONFIGS = {
    "cfg1": dict(views=20, channels=64, height=120, width=160, voxel_dim=(40, 40, 16), voxel_size=0.16,
                 tsdf="random", grids=300, dtype="f32"),
    "cfg2": dict(views=50, channels=256, height=120, width=160, voxel_dim=(80, 80, 32), voxel_size=0.08,
                 tsdf="room", grids=300, dtype="f32"),
    "cfg3": dict(views=100, channels=256, height=192, width=256, voxel_dim=(96, 96, 40), voxel_size=0.08,
                 tsdf="room", grids=300, dtype="f32"),
    "cfg4": dict(views=50, channels=256, height=120, width=160, voxel_dim=(160, 160, 64), voxel_size=0.04,
                 tsdf="room", grids=300, dtype="f32"),
    "cfg4_c32": dict(views=50, channels=32, height=120, width=160, voxel_dim=(160, 160, 64), voxel_size=0.04,
                     tsdf="room", grids=300, dtype="f32"),
    "cfg5": dict(views=300, channels=128, height=120, width=160, voxel_dim=(256, 256, 96), voxel_size=0.04,
                 tsdf="room", grids=256, dtype="bf16"),
    # the reference's own shipped test configuration (ray_marching_scannet.py:12-19, :123, :158): 32 channels
    "ref_test": dict(views=50, channels=32, height=120, width=160, voxel_dim=(256, 256, 96), voxel_size=0.04,
                     tsdf="room", grids=300, dtype="f32"),
    # small shapes for fast parity tests
    "tiny": dict(views=3, channels=8, height=12, width=16, voxel_dim=(8, 8, 4), voxel_size=0.4,
                 tsdf="room", grids=40, dtype="f32"),
    "small": dict(views=6, channels=16, height=30, width=40, voxel_dim=(20, 20, 8), voxel_size=0.25,
                  tsdf="room", grids=100, dtype="f32"),
    # odd grid sizes (border handling of the march's distance field and of partially filled warps)
    "odd": dict(views=4, channels=8, height=20, width=28, voxel_dim=(13, 10, 7), voxel_size=0.3,
                tsdf="room", grids=120, dtype="f32"),
    # a room with large regions of constant TSDF (exercises the march's empty-space jumps), full-length march
    "room40": dict(views=4, channels=8, height=30, width=40, voxel_dim=(40, 40, 16), voxel_size=0.16,
                   tsdf="room", grids=300, dtype="f32"),
}


@dataclass
class Scene:
    """One synthetic scene.  Arrays are numpy, float32 unless noted."""
    name: str
    seed: int
    voxel_dim: tuple
    voxel_size: float
    origin: np.ndarray            # [3] float32
    stride: int
    grids: int
    projections: np.ndarray       # [V, 3, 4] un-scaled (full-resolution intrinsics)
    features: np.ndarray = None   # [V, C, H, W] (logical NCHW) or None when generated on device
    tsdf: np.ndarray = None       # [nx, ny, nz]
    height: int = 0
    width: int = 0
    channels: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def views(self):
        return self.projections.shape[0]

    @property
    def nvox(self):
        nx, ny, nz = self.voxel_dim
        return nx * ny * nz

    @property
    def voxel_views(self):
        return self.views * self.nvox

    @property
    def ray_steps(self):
        return self.views * self.height * self.width * self.grids


def ring_cameras(views, height, width, stride, extent, rng, return_poses=False):
    """Pinhole cameras on a ring inside the room, looking across it.

    fx = fy = 0.9 * W_img (ScanNet's 577/640), principal point at the image
    centre, W_img = stride * W.  Returns projections [V,3,4] float32 =
    K @ inv(E)[:3] evaluated in float64 and rounded once (and, on request, K and the poses E).
    """
    w_img, h_img = stride * width, stride * height
    k = np.array([[0.9 * w_img, 0.0, 0.5 * w_img],
                  [0.0, 0.9 * w_img, 0.5 * h_img],
                  [0.0, 0.0, 1.0]], dtype=np.float64)
    ex, ey, ez = extent
    centre = np.array([0.5 * ex, 0.5 * ey, 0.45 * ez])
    radius = 0.22 * min(ex, ey)
    out = np.empty((views, 3, 4), dtype=np.float32)
    poses = np.empty((views, 4, 4), dtype=np.float64)
    for i in range(views):
        ang = 2.0 * math.pi * (i + 0.25 * rng.random()) / views
        cam = centre + np.array([radius * math.cos(ang), radius * math.sin(ang), 0.1 * ez * (rng.random() - 0.5)])
        pitch = math.radians(20.0 * (rng.random() - 0.5))
        yaw = ang + math.pi + math.radians(30.0 * (rng.random() - 0.5))   # look across the room
        fwd = np.array([math.cos(yaw) * math.cos(pitch), math.sin(yaw) * math.cos(pitch), math.sin(pitch)])
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        pose = np.eye(4)
        pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, down, fwd, cam
        poses[i] = pose
        out[i] = (k @ np.linalg.inv(pose)[:3, :]).astype(np.float32)
    return (out, k, poses) if return_poses else out


def room_depth_maps(k, poses, height, width, extent, rng=None, holes=0.02):
    """Depth maps [V,H,W] (metres along the optical axis, float32) of the box room whose walls sit at 15 % / 85 % of
    the extent, seen from cameras inside it; a fraction `holes` of the pixels is set to 0 (no depth reading), the
    value the fusion code treats as invalid (data_prepare/scannet/tsdf.py:422-423)."""
    lo = np.array([0.15 * e for e in extent])
    hi = np.array([0.85 * e for e in extent])
    u, v = np.meshgrid(np.arange(width), np.arange(height))
    rays_cam = np.stack([(u - k[0, 2]) / k[0, 0], (v - k[1, 2]) / k[1, 1], np.ones_like(u, dtype=np.float64)], axis=-1)
    out = np.empty((len(poses), height, width), dtype=np.float32)
    for i, pose in enumerate(poses):
        d = rays_cam @ pose[:3, :3].T                      # world directions, z_cam = 1 per unit of t
        c = pose[:3, 3]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(d > 0, (hi - c) / d, np.where(d < 0, (lo - c) / d, np.inf))
        depth = t.min(axis=-1)
        if rng is not None and holes > 0:
            depth = np.where(rng.random(depth.shape) < holes, 0.0, depth)
        out[i] = depth.astype(np.float32)
    return out


def room_tsdf(voxel_dim, trunc_voxels=3.0, walls=(0.15, 0.85)):
    """Box room: walls at 15 % / 85 % of the extent on every axis, truncation 3 voxels."""
    nx, ny, nz = voxel_dim
    gx, gy, gz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    d = np.full(voxel_dim, np.inf)
    for g, n in ((gx, nx), (gy, ny), (gz, nz)):
        lo, hi = walls[0] * (n - 1), walls[1] * (n - 1)
        d = np.minimum(d, np.minimum(g - lo, hi - g))      # >0 inside the room, in voxels
    sdf = -d / trunc_voxels                                # free space negative, behind walls positive
    tsdf = np.clip(sdf, -1.0, 1.0)
    tsdf[sdf >= 1.0] = 1.0                                  # beyond truncation behind the wall: unobserved
    return (tsdf * 1.05 * 0.95).astype(np.float32)         # stay inside tanh*1.05's range


def make_scene(config="cfg2", seed=0, with_features=True, stride=4, **override):
    """Builds scene `seed` of a named configuration (or a dict of the same keys)."""
    cfg = dict(CONFIGS[config]) if isinstance(config, str) else dict(config)
    cfg.update(override)
    name = config if isinstance(config, str) else cfg.get("name", "custom")
    rng = np.random.default_rng(seed)
    nx, ny, nz = cfg["voxel_dim"]
    vs = float(cfg["voxel_size"])
    extent = (nx * vs, ny * vs, nz * vs)
    proj = ring_cameras(cfg["views"], cfg["height"], cfg["width"], stride, extent, rng)
    if cfg["tsdf"] == "random":
        tsdf = rng.uniform(-1.0, 1.0, size=(nx, ny, nz)).astype(np.float32)
    elif cfg["tsdf"] == "room_var":
        # the room, with seed-dependent wall positions and truncation width, so that the number of samples the
        # march keeps (M) differs from scene to scene; its own generator: the other draws are those of "room"
        r2 = np.random.default_rng(7777 + seed)
        tsdf = room_tsdf((nx, ny, nz), trunc_voxels=float(r2.uniform(2.0, 4.5)),
                         walls=(float(r2.uniform(0.10, 0.20)), float(r2.uniform(0.80, 0.90))))
    else:
        tsdf = room_tsdf((nx, ny, nz))
    feats = None
    if with_features:
        feats = rng.standard_normal((cfg["views"], cfg["channels"], cfg["height"], cfg["width"]),
                                    dtype=np.float32)
    return Scene(name=name, seed=seed, voxel_dim=(nx, ny, nz), voxel_size=vs,
                 origin=np.zeros(3, dtype=np.float32), stride=stride, grids=int(cfg["grids"]),
                 projections=proj, features=feats, tsdf=tsdf, height=cfg["height"], width=cfg["width"],
                 channels=cfg["channels"], meta=dict(cfg))


def device_features(scene, device, dtype=None, channels_last=True):
    """Bench-sized feature stack generated on `device` (seeded by the scene's seed).

    Returns a tensor of logical shape [V, 1, C, H, W]; with `channels_last` the
    physical layout is [V, 1, H, W, C] (torch's channels_last on the trailing
    NCHW dims), the layout the gather kernels read natively.
    """
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(1000 + scene.seed)
    v, c, h, w = scene.views, scene.channels, scene.height, scene.width
    dtype = dtype or torch.float32
    if channels_last:
        x = torch.randn((v, 1, h, w, c), generator=g, device=device, dtype=torch.float32).to(dtype)
        return x.permute(0, 1, 4, 2, 3)
    return torch.randn((v, 1, c, h, w), generator=g, device=device, dtype=torch.float32).to(dtype)
