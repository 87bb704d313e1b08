"""TSDF head hand-over (projects/mvsdetection/models/atlas_head.py:15-82 `AtlasTSDFHead`):
oracle and CUDA path against outputs, losses and gradients of the unmodified reference class (tests/golden_head,
oracle/make_golden_head.py).

This is a floating-point row: the channel reduction order of the reference's convolution is unspecified and tanh is
a libm call, so values are compared to 1e-5 (absolute, on a range of +-1.05).  The surface mask |prev| < thr is exact
when the previous scale is given bit-identically; in chained runs it may differ only where |prev| is within 1e-5 of
the threshold (band protocol, as for the NeuS weights)."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden_head", "*.npz")))
TOL = 1e-5
BAND = 1e-5


def _load(path):
    with np.load(path) as z:
        g = {k: z[k] for k in z.files}
    g["ls"] = float(g["label_smoothing"])
    g["thr"] = [float(t) for t in g["sparse_threshold"]]
    g["keys"] = [str(k) for k in g["keys"]]
    return g


@pytest.fixture(params=CASES, ids=lambda p: os.path.basename(p)[:-4])
def head(request):
    return _load(request.param)


def _upsampled(prev):
    return np.repeat(np.repeat(np.repeat(prev, 2, axis=-3), 2, axis=-2), 2, axis=-1)


def _outside_band(prev, thr):
    return np.abs(np.abs(_upsampled(prev)) - thr) > BAND


def test_oracle_matches_reference_head_per_scale(head):
    g = head
    for i in range(3):
        prev = g[f"tsdf{i - 1}"][0, 0] if i > 0 else None
        t, m = oracle.tsdf_head_scale(g[f"x{i}"][0], g[f"w{i}"], prev, g["ls"], g["thr"][i - 1] if i > 0 else None)
        assert np.max(np.abs(t - g[f"tsdf{i}"][0, 0])) <= TOL
        if i > 0:                                   # same prev bits in, same mask out
            assert np.array_equal(m, np.abs(_upsampled(prev)) < np.float32(g["thr"][i - 1]))
            assert np.array_equal(t[~m], (np.sign(_upsampled(prev)) * np.float32(0.999))[~m])


def test_oracle_matches_reference_head_chained(head):
    g = head
    ts, ms = oracle.tsdf_head([g[f"x{i}"][0] for i in range(3)], [g[f"w{i}"] for i in range(3)], g["ls"], g["thr"])
    for i in range(3):
        ok = np.ones(ts[i].shape, bool) if i == 0 else _outside_band(g[f"tsdf{i - 1}"][0, 0], g["thr"][i - 1])
        assert ok.mean() > 0.99
        assert np.max(np.abs(ts[i] - g[f"tsdf{i}"][0, 0])[ok]) <= TOL


def test_mirror_parameters_and_losses_match_reference(head):
    """State-dict names / shapes of the reference class, and its loss bookkeeping (ah.py:57-82) on the golden outputs."""
    import cnrma_b200 as cn
    g = head
    chans = [int(c) for c in g["channels"]]
    m = cn.AtlasTSDFHead(chans, 3, 0.04, g["ls"], g["thr"])
    assert m.keys == g["keys"]
    sd = m.state_dict()
    assert list(sd) == ["decoders.0.weight", "decoders.1.weight", "decoders.2.weight"]
    assert [tuple(v.shape) for v in sd.values()] == [(1, c, 1, 1, 1) for c in chans[::-1]]
    output = {"scene_tsdf_" + k: torch.from_numpy(g[f"tsdf{i}"]) for i, k in enumerate(g["keys"])}
    targets = {"tsdf_gt_" + k: torch.from_numpy(g[f"target{i}"]) for i, k in enumerate(g["keys"])}
    masks = [torch.from_numpy(np.abs(_upsampled(g[f"tsdf{i - 1}"])) < np.float32(g["thr"][i - 1])) for i in (1, 2)]
    losses = m.losses(output, masks, targets)
    for i, k in enumerate(g["keys"]):
        assert abs(losses["tsdf_loss_" + k].item() - float(g[f"loss{i}"])) <= 1e-6


def test_head_needs_cuda():
    import cnrma_b200 as cn
    with pytest.raises(cn.CnrmaError):
        cn.tsdf_head_scale(torch.zeros(1, 4, 4, 4, 4), torch.zeros(1, 4, 1, 1, 1))


# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cn():
    import cnrma_b200
    cnrma_b200.load()
    return cnrma_b200


def _layouts(x):
    yield "ncdhw", x
    yield "channels_last_3d", x.contiguous(memory_format=torch.channels_last_3d)
    yield "bf16", x.to(torch.bfloat16)


@pytest.mark.gpu
def test_gpu_head_scale_vs_reference_and_oracle(cn, head):
    g = head
    for i in range(3):
        prev = torch.from_numpy(g[f"tsdf{i - 1}"]).cuda() if i > 0 else None
        thr = g["thr"][i - 1] if i > 0 else 0.0
        w = torch.from_numpy(g[f"w{i}"]).cuda().view(1, -1, 1, 1, 1)
        for name, x in _layouts(torch.from_numpy(g[f"x{i}"]).cuda()):
            t, m = cn.tsdf_head_scale(x, w, prev, g["ls"], thr)
            if name == "bf16":
                ref, om = oracle.tsdf_head_scale(x.float().cpu().numpy()[0], g[f"w{i}"],
                                                 None if prev is None else g[f"tsdf{i - 1}"][0, 0], g["ls"], thr if i else None)
            else:
                ref, om = g[f"tsdf{i}"][0, 0], None
                o2, om = oracle.tsdf_head_scale(g[f"x{i}"][0], g[f"w{i}"], None if prev is None else g[f"tsdf{i - 1}"][0, 0],
                                                g["ls"], thr if i else None)
                assert np.max(np.abs(t[0, 0].cpu().numpy() - o2)) <= TOL, (name, i)
            assert np.max(np.abs(t[0, 0].cpu().numpy() - ref)) <= TOL, (name, i)
            assert np.array_equal(m[0, 0].cpu().numpy(), om), (name, i)
            if i > 0:                                # the sparsified voxels are exact
                up = _upsampled(g[f"tsdf{i - 1}"][0, 0])
                assert np.array_equal(t[0, 0].cpu().numpy()[~om], (np.sign(up) * np.float32(0.999))[~om])


@pytest.mark.gpu
def test_gpu_head_forward_losses_and_gradients_vs_reference(cn, head):
    """The mirror class end to end: forward(xs, targets) then backward of the summed losses, against the reference's
    outputs, losses, d/dx and d/dweight (float32, 1e-5 of each tensor's range)."""
    g = head
    chans = [int(c) for c in g["channels"]]
    m = cn.AtlasTSDFHead(chans, 3, 0.04, g["ls"], g["thr"]).cuda()
    m.load_state_dict({f"decoders.{i}.weight": torch.from_numpy(g[f"w{i}"]).view(1, -1, 1, 1, 1) for i in range(3)})
    for layout in ("ncdhw", "channels_last_3d"):
        m.zero_grad()
        xs = []
        for i in range(3):
            x = torch.from_numpy(g[f"x{i}"]).cuda()
            if layout == "channels_last_3d":
                x = x.contiguous(memory_format=torch.channels_last_3d)
            xs.append(x.requires_grad_(True))
        targets = {"tsdf_gt_" + k: torch.from_numpy(g[f"target{i}"]).cuda() for i, k in enumerate(g["keys"])}
        out, losses = m(xs, targets)
        sum(losses.values()).backward()
        for i, k in enumerate(g["keys"]):
            ok = np.ones(g[f"tsdf{i}"].shape, bool) if i == 0 else _outside_band(g[f"tsdf{i - 1}"], g["thr"][i - 1])
            assert np.max(np.abs(out["scene_tsdf_" + k].detach().cpu().numpy() - g[f"tsdf{i}"])[ok]) <= TOL
            assert abs(losses["tsdf_loss_" + k].item() - float(g[f"loss{i}"])) <= 1e-5
            gx, gw = xs[i].grad.cpu().numpy(), m.decoders[i].weight.grad.cpu().numpy().reshape(-1)
            assert np.max(np.abs(gx - g[f"grad_x{i}"])) <= TOL * max(1e-30, np.max(np.abs(g[f"grad_x{i}"]))), (layout, i)
            assert np.max(np.abs(gw - g[f"grad_w{i}"])) <= TOL * max(1e-30, np.max(np.abs(g[f"grad_w{i}"]))), (layout, i)


@pytest.mark.gpu
def test_gpu_head_inference_feeds_the_march(cn):
    """no_grad forward on a larger random volume vs the plain PyTorch fp32 formulation of ah.py:38-52, and the finest
    output goes straight into the ray march."""
    from cnrma_b200 import synthetic as S
    torch.manual_seed(5)
    torch.backends.cudnn.allow_tf32 = False          # compare against fp32 PyTorch, not cudnn's TF32 default
    chans, fine = [8, 16, 32], (40, 40, 16)
    m = cn.AtlasTSDFHead(chans, 3, 0.16, 1.05, [0.99, 0.99, 0.99]).cuda()
    xs = [3 * torch.randn((1, c) + tuple(d // 2 ** (2 - i) for d in fine), device="cuda") for i, c in enumerate(chans[::-1])]
    with torch.no_grad():
        out, _ = m(xs)
        prev = None
        for i, (dec, x) in enumerate(zip(m.decoders, xs)):
            t = torch.tanh(dec(x)) * 1.05
            if i > 0:
                up = torch.nn.functional.interpolate(prev, scale_factor=2)
                keep = up.abs() < 0.99
                t[~keep] = up[~keep].sign() * .999
            mine = out["scene_tsdf_" + m.keys[i]]
            ok = torch.ones_like(t, dtype=torch.bool) if i == 0 else (up.abs() - 0.99).abs() > BAND
            assert float((mine - t).abs()[ok].max()) <= TOL
            prev = mine                                  # chain on our own output so that bands do not accumulate
    sc = S.make_scene("cfg1")
    assert tuple(sc.voxel_dim) == fine
    p = torch.from_numpy(sc.projections).cuda().unsqueeze(1)
    f = torch.from_numpy(sc.features).cuda().unsqueeze(1)
    pts = cn.rma_points(p, f, out["scene_tsdf_" + m.keys[-1]], sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                        threshold=0.05)
    assert pts[0] is None or pts[0].shape[1] == 3 + sc.channels
