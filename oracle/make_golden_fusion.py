"""TEST INFRASTRUCTURE ONLY -- golden vectors for the GT TSDF fusion row, produced by
the UNMODIFIED reference class `TSDFFusion` of data_prepare/scannet/tsdf.py (imported from /root/reference with
matplotlib / skimage / trimesh stubbed).  Run in the build container:  python oracle/make_golden_fusion.py"""
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "cn-rma_b200"))
import synthetic  # noqa: E402

REF = os.environ.get("CNRMA_REFERENCE_ROOT", "/root/reference")


def load_reference_fusion():
    for name in ("matplotlib", "matplotlib.cm", "skimage", "skimage.measure", "trimesh"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            m.__getattr__ = lambda attr, _n=name: mock.MagicMock(name=f"{_n}.{attr}")
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_scannet_tsdf", os.path.join(REF, "data_prepare", "scannet", "tsdf.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_fusion()
    for name, dim, vs, frames, hw, seed in (("fusion_small", (24, 24, 10), 0.25, 6, (48, 64), 1),
                                            ("fusion_odd", (19, 14, 9), 0.3, 4, (30, 40), 2)):
        rng = np.random.default_rng(seed)
        extent = tuple(d * vs for d in dim)
        origin = (0.0, 0.0, 0.0) if name == "fusion_small" else (-0.2, 0.1, 0.05)
        P, k, poses = synthetic.ring_cameras(frames, hw[0], hw[1], 1, extent, rng, return_poses=True)
        depth = synthetic.room_depth_maps(k, poses, hw[0], hw[1], extent, rng)
        color = rng.uniform(0, 255, size=(frames, 3) + hw).astype(np.float32)
        label = rng.integers(0, 40, size=(frames,) + hw).astype(np.int64)
        fus = ref.TSDFFusion(dim, vs, origin, trunc_ratio=3, device=torch.device("cpu"), color=True, label=True)
        for i in range(frames):
            fus.integrate(torch.from_numpy(P[i]), torch.from_numpy(depth[i]), torch.from_numpy(color[i]),
                          torch.from_numpy(label[i]))
        out = dict(voxel_dim=np.array(dim), voxel_size=np.float64(vs), origin=np.asarray(origin, np.float32),
                   trunc_ratio=np.float64(3), projections=P, depth=depth, color_img=color, label_img=label,
                   tsdf_vol=fus.tsdf_vol.numpy().copy(), weight_vol=fus.weight_vol.numpy().copy(),
                   color_vol=fus.color_vol.numpy().copy(), label_vol=fus.label_vol.numpy().copy())
        # get_tsdf(): the normalisation lines (tsdf.py:460-470), without the skimage/trimesh container
        t = fus.tsdf_vol.clone()
        t[fus.weight_vol > 0] /= fus.weight_vol[fus.weight_vol > 0]
        c = fus.color_vol.clone()
        c[:, fus.weight_vol > 0] /= fus.weight_vol[fus.weight_vol > 0]
        out["tsdf_final"] = t.numpy()
        out["color_final"] = c.numpy()
        path = os.path.join(ROOT, "tests", "golden_fusion", name + ".npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **out)
        print(name, "observed voxels", int((fus.weight_vol > 0).sum()), "of", fus.weight_vol.numel(), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
