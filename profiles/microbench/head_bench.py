"""TSDF head (atlas_head.py:38-52) at the reference's test grid: xs = [128ch @ 64x64x24, 64ch @ 128x128x48,
32ch @ 256x256x96]; fused kernels vs the plain PyTorch formulation of the same lines, forward (no_grad) and
forward+backward.  Two input regimes: 'dense' (random features: ~40 % of the fine volume stays surface) and 'room'
(features built so that the head predicts a room TSDF: a few % surface, the regime of real scenes)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cnrma_b200 as cn
from cnrma_b200 import synthetic

fine, chans = (256, 256, 96), [32, 64, 128]
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False      # the comparison is against fp32 PyTorch (cudnn would use TF32 by default)
head = cn.AtlasTSDFHead(chans, 3, 0.04, 1.05, [0.99, 0.99, 0.99]).cuda()


def torch_forward(xs):
    out, prev = [], None
    for i, (dec, x) in enumerate(zip(head.decoders, xs)):
        t = torch.tanh(dec(x)) * 1.05
        if i > 0:
            up = F.interpolate(prev, scale_factor=2)
            keep = up.abs() < 0.99
            t[~keep] = up[~keep].sign() * .999
        out.append(t)
        prev = t
    return out


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


def make_inputs(kind):
    xs = []
    for i, c in enumerate(chans[::-1]):
        dims = tuple(d // 2 ** (2 - i) for d in fine)
        x = torch.randn((1, c) + dims, device="cuda")
        if kind == "room":
            # channel 0 carries atanh(tsdf/1.05) / w0 of a room TSDF at this scale, the others small noise
            t = torch.from_numpy(synthetic.room_tsdf(dims)).cuda().clamp(-1.04, 1.04)
            w = head.decoders[i].weight.detach().view(-1)
            x *= 0.01
            x[0, 0] = torch.atanh(t / 1.05) / w[0]
        xs.append(x)
    return xs


for kind in ("dense", "room"):
    xs = make_inputs(kind)
    with torch.no_grad():
        out, _ = head(xs)
        ref = torch_forward(xs)
        surf = [float((o.abs() < 0.99).float().mean()) for o in out.values()]
        err = max(float((a - b).abs().max()) for a, b in zip(out.values(), ref))
        t_mine = timed(lambda: head(xs), 50)
        t_torch = timed(lambda: torch_forward(xs), 20)
    in_bytes = sum(x.numel() * 4 for x in xs)
    print(f"[{kind}] surface fraction per scale {['%.3f' % s for s in surf]}  max |fused - torch| = {err:.2e}")
    print(f"[{kind}] forward: fused {t_mine:.3f} ms  torch {t_torch:.3f} ms  ({t_torch / t_mine:.1f}x); "
          f"dense input bytes {in_bytes / 1e6:.0f} MB -> {in_bytes / t_mine / 1e6:.0f} GB/s equivalent")
    xg = [x.clone().requires_grad_(True) for x in xs]

    def step_mine():
        head.zero_grad(set_to_none=True)
        for x in xg:
            x.grad = None
        o, _ = head(xg)
        sum(v.square().mean() for v in o.values()).backward()

    def step_torch():
        head.zero_grad(set_to_none=True)
        for x in xg:
            x.grad = None
        sum(v.square().mean() for v in torch_forward(xg)).backward()

    def fwd_mine():
        head(xg)

    def fwd_torch():
        torch_forward(xg)

    f_mine, f_torch = timed(fwd_mine, 5), timed(fwd_torch, 5)
    t_mine = timed(step_mine, 10)
    gm = [x.grad.clone() for x in xg] + [d.weight.grad.clone() for d in head.decoders]
    t_torch = timed(step_torch, 5)
    gt = [x.grad.clone() for x in xg] + [d.weight.grad.clone() for d in head.decoders]
    torch.backends.cudnn.allow_tf32 = True
    t_torch_tf32, f_torch_tf32 = timed(step_torch, 5), timed(fwd_torch, 5)
    torch.backends.cudnn.allow_tf32 = False
    rel = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(gm, gt))
    print(f"[{kind}] forward+backward: fused {t_mine:.3f} ms (forward with graph {f_mine:.3f})  torch {t_torch:.3f} ms "
          f"(forward {f_torch:.3f}); torch with cudnn TF32 (its default) {t_torch_tf32:.3f} ms (forward {f_torch_tf32:.3f}); "
          f"max rel grad diff vs fp32 torch {rel:.2e}")
