"""GT TSDF fusion (data_prepare/scannet/tsdf.py:353-475 `TSDFFusion`): the oracle and the
CUDA path against volumes produced by the unmodified reference class (tests/golden_fusion, oracle/make_golden_fusion.py).
Everything is bit-exact: the per-voxel update order is the frame order in both."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden_fusion", "*.npz")))


def _load(path):
    with np.load(path) as z:
        g = {k: z[k] for k in z.files}
    g["voxel_dim"] = tuple(int(v) for v in g["voxel_dim"])
    g["voxel_size"] = float(g["voxel_size"])
    g["trunc_margin"] = g["voxel_size"] * float(g["trunc_ratio"])      # tsdf.py:373
    return g


@pytest.fixture(params=CASES, ids=lambda p: os.path.basename(p)[:-4])
def fusion(request):
    return _load(request.param)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_oracle_matches_reference_fusion(fusion):
    g = fusion
    n = int(np.prod(g["voxel_dim"]))
    tsdf = np.ones(n, np.float32)
    weight = np.zeros(n, np.float32)
    color = np.zeros((3, n), np.float32)
    label = -np.ones(n, np.int64)
    for i in range(g["projections"].shape[0]):
        oracle.tsdf_integrate(g["voxel_dim"], g["voxel_size"], g["origin"], g["projections"][i], g["depth"][i],
                              g["trunc_margin"], tsdf, weight, g["color_img"][i], color, g["label_img"][i], label)
    assert np.array_equal(_bits(weight), _bits(g["weight_vol"]))
    assert np.array_equal(_bits(tsdf), _bits(g["tsdf_vol"]))
    assert np.array_equal(_bits(color), _bits(g["color_vol"]))
    assert np.array_equal(label, g["label_vol"])


@pytest.mark.gpu
def test_cuda_fusion_matches_reference(fusion):
    import cnrma_b200 as cn
    g = fusion
    fus = cn.TSDFFusion(g["voxel_dim"], g["voxel_size"], g["origin"], trunc_ratio=float(g["trunc_ratio"]),
                        device="cuda", color=True, label=True)
    P = torch.from_numpy(g["projections"]).cuda()
    D = torch.from_numpy(g["depth"]).cuda()
    Cimg = torch.from_numpy(g["color_img"]).cuda()
    L = torch.from_numpy(g["label_img"]).cuda()
    half = P.shape[0] // 2
    for i in range(half):                                    # frame by frame, like generate_tsdf.py:108-121
        fus.integrate(P[i], D[i], Cimg[i], L[i])
    fus.integrate_frames(P[half:], D[half:], Cimg[half:], L[half:])     # and the rest in one launch
    assert np.array_equal(_bits(fus.weight_vol.cpu().numpy()), _bits(g["weight_vol"]))
    assert np.array_equal(_bits(fus.tsdf_vol.cpu().numpy()), _bits(g["tsdf_vol"]))
    assert np.array_equal(_bits(fus.color_vol.cpu().numpy()), _bits(g["color_vol"]))
    assert np.array_equal(fus.label_vol.cpu().numpy(), g["label_vol"])
    tsdf, attrs = fus.get_tsdf()
    assert np.array_equal(_bits(tsdf.cpu().numpy().reshape(-1)), _bits(g["tsdf_final"]))
    assert np.array_equal(_bits(attrs["color"].cpu().numpy().reshape(3, -1)), _bits(g["color_final"]))
    fus.reset()
    assert float(fus.weight_vol.abs().max()) == 0.0 and float(fus.tsdf_vol.min()) == 1.0
