"""TEST INFRASTRUCTURE ONLY -- golden vectors for the TSDF head row, produced by the
UNMODIFIED reference class `AtlasTSDFHead` (projects/mvsdetection/models/atlas_head.py, imported from
/root/reference through oracle/ref_shim.py).  The reference's loss code calls `.cuda()` on a zero scalar
(atlas_head.py:60); for this CPU run `torch.Tensor.cuda` is patched to the identity -- torch is patched, the
reference file is not.  Run in the build container:  python oracle/make_golden_head.py"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def main():
    ref_shim.load_reference()
    head_mod = importlib.import_module("projects.mvsdetection.models.atlas_head")
    torch.Tensor.cuda = lambda self, *a, **k: self
    os.makedirs(os.path.join(ROOT, "tests", "golden_head"), exist_ok=True)
    for name, chans, fine, thr, seed in (("head_small", [8, 16, 32], (16, 16, 8), [0.99, 0.99, 0.99], 11),
                                         ("head_odd", [4, 12, 20], (24, 12, 20), [0.9, 0.8, 0.99], 12)):
        torch.manual_seed(seed)
        head = head_mod.AtlasTSDFHead(chans, 3, 0.04, 1.05, thr)
        xs = []
        for i, c in enumerate(chans[::-1]):
            f = 2 ** (2 - i)
            dims = tuple(d // f for d in fine)
            xs.append((3.0 * torch.randn((1, c) + dims)).requires_grad_(True))
        targets = {}
        for i, key in enumerate(head.keys):
            f = 2 ** (2 - i)
            dims = tuple(d // f for d in fine)
            t = torch.empty((1, 1) + dims).uniform_(-1, 1)
            t[torch.rand((1, 1) + dims) < 0.3] = 1.0
            t[:, :, : dims[0] // 4] = 1.0                     # whole z-columns "outside" (atlas_head.py:69)
            targets["tsdf_gt_" + key] = t
        out, losses = head(xs, targets)
        total = sum(losses.values())
        total.backward()
        rec = dict(channels=np.array(chans), fine_dim=np.array(fine), label_smoothing=np.float64(1.05),
                   sparse_threshold=np.array(thr, np.float64), keys=np.array(head.keys))
        for i, key in enumerate(head.keys):
            rec[f"x{i}"] = xs[i].detach().numpy()
            rec[f"w{i}"] = head.decoders[i].weight.detach().numpy().reshape(-1)
            rec[f"tsdf{i}"] = out["scene_tsdf_" + key].detach().numpy()
            rec[f"target{i}"] = targets["tsdf_gt_" + key].numpy()
            rec[f"loss{i}"] = np.float64(losses["tsdf_loss_" + key].item())
            rec[f"grad_x{i}"] = xs[i].grad.numpy()
            rec[f"grad_w{i}"] = head.decoders[i].weight.grad.numpy().reshape(-1)
        path = os.path.join(ROOT, "tests", "golden_head", name + ".npz")
        np.savez_compressed(path, **rec)
        print(path, {k: float(v) for k, v in losses.items()},
              [float((np.abs(rec[f"tsdf{i}"]) >= 0.998).mean()) for i in range(3)])


if __name__ == "__main__":
    main()
