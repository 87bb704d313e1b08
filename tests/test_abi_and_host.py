"""CPU: the C-ABI library loads and exports every symbol include/cnrma_b200.h declares (no compute calls),
ctypes struct layouts match the header, and the host-side logic behaves."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import cnrma_b200 as cn
from cnrma_b200 import _lib, distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "cnrma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cnrma_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    cn.build()
    lib = cn.load()
    declared = _declared_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cnrma_b200.h but not exported"
    assert sorted(cn.EXPORTS) == declared, "ctypes signature table and header disagree"
    assert lib.cnrma_abi_version() == 2
    assert lib.cnrma_status_string(-2).decode().startswith("feature maps")


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Grid) == 28
    assert C.sizeof(_lib.RmaResult) == 24 or C.sizeof(_lib.RmaResult) == 32
    assert _lib.Features.view_ptrs_host.offset == 48
    assert _lib.RmaResult.weight_sum.offset == 8 and _lib.RmaResult.mean.offset == 16


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sc = cn.synthetic.make_scene("tiny", seed=0)
    f = torch.from_numpy(sc.features).unsqueeze(1)
    p = torch.from_numpy(sc.projections).unsqueeze(1)
    with pytest.raises(cn.CnrmaError):
        cn.aggregate_views(p, f, sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride)
    with pytest.raises(cn.CnrmaError):
        cn.rma_points(p, f, torch.from_numpy(sc.tsdf)[None, None], sc.voxel_dim, sc.voxel_size, sc.origin, sc.stride,
                      threshold=0.05)


def test_t_one_matches_reference_formula():
    import math
    lib = cn.load()
    for dim, vs, n in [((80, 80, 32), 0.08, 300), ((256, 256, 96), 0.04, 300), ((8, 8, 4), 0.4, 40)]:
        g = _lib.make_grid(dim, vs, (0, 0, 0))
        want = np.float32(math.sqrt(dim[0] ** 2 + dim[1] ** 2 + dim[2] ** 2) * vs / n)   # rm.py:710-711
        assert np.float32(lib.cnrma_t_one(C.byref(g), vs, n)) == want


def test_workspace_query():
    lib = cn.load()
    n = C.c_size_t(0)
    g = _lib.make_grid((80, 80, 32), 0.08, (0, 0, 0))
    assert lib.cnrma_rma_workspace_bytes(C.byref(g), 50, 120, 160, 300, 0, 0.05, 0, C.byref(n)) == 0
    rays = 50 * 120 * 160
    assert n.value >= rays * 4 + 2 * rays * 21 * 4          # counts + (step, weight) records, 1/0.05 + 1 per ray
    assert lib.cnrma_rma_workspace_bytes(C.byref(g), 0, 120, 160, 300, 0, 0.05, 0, C.byref(n)) == -1
    assert lib.cnrma_rma_workspace_bytes(C.byref(g), 1, 1, 1, 1, 7, 0.05, 0, C.byref(n)) == -1


def test_scale_and_invert_projections_follow_reference_ops():
    sc = cn.synthetic.make_scene("tiny", seed=3)
    p = torch.from_numpy(sc.projections)
    ps = cn.scale_projections(p, 4)
    want = p.clone()
    want[:, :2, :] = want[:, :2, :] / 4                         # rm.py:238-239
    assert torch.equal(ps, want)
    inv = cn.invert_projections(ps)
    for v in range(ps.shape[0]):                                # one torch.inverse per matrix, rm.py:96-102
        p4 = torch.cat((ps[v], torch.tensor([[0.0, 0, 0, 1]])), 0)
        assert torch.equal(inv[v], torch.inverse(p4))


def test_shards():
    for v, w in [(50, 8), (50, 4), (7, 2), (3, 8)]:
        ranges = [D.view_shard(v, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == v
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1
    assert D.scene_shard(10, 1, 4) == [1, 5, 9]


def test_synthetic_scene_shapes():
    sc = cn.synthetic.make_scene("cfg2", seed=0, with_features=False)
    assert sc.voxel_views == 10_240_000 and sc.ray_steps == 288_000_000
    assert sc.projections.shape == (50, 3, 4) and sc.tsdf.shape == (80, 80, 32)
    assert np.abs(sc.tsdf).max() <= 1.05


def test_box_shards_tile_the_grid():
    for dim, w in [((160, 160, 64), 2), ((160, 160, 64), 4), ((160, 160, 64), 8), ((20, 20, 8), 8), ((13, 10, 7), 3),
                   ((13, 10, 7), 6), ((80, 80, 32), 16), ((5, 4, 3), 1)]:
        seen = np.zeros(dim, np.int32)
        for r in range(w):
            lo, bd = D.box_shard(dim, r, w)
            seen[tuple(slice(l, l + d) for l, d in zip(lo, bd))] += 1
        assert (seen == 1).all(), (dim, w)
        gx, gy, gz = D.default_splits(w)
        assert gx * gy * gz == w
    assert D.default_splits(2) == (1, 1, 2) and D.default_splits(4) == (2, 1, 2) and D.default_splits(8) == (2, 2, 2)
    with pytest.raises(ValueError):
        D.box_shard((4, 4, 1), 0, 2)                  # the default cut for two ranks is along z
    assert D.box_shard((4, 4, 1), 1, 2, splits=(2, 1, 1)) == ((2, 0, 0), (2, 4, 1))
    for nx, parts in [(160, 4), (7, 3), (3, 5)]:
        ch = D.x_chunks(nx, parts)
        assert ch[0][0] == 0 and ch[-1][1] == nx and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        assert all(b > a for a, b in ch)


def test_box_struct_layout_and_reference_staging():
    assert C.sizeof(_lib.Box) == 24 and _lib.Box.dim.offset == 12
    import build_ref
    import ref_shim
    # wherever the reference tree is mounted the staged copy is complete; elsewhere the stage step is a no-op
    if os.path.isdir("/root/reference/projects/mvsdetection"):
        assert build_ref.stage()
        assert os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "projects", "mvsdetection", "models", "ray_marching.py"))
    assert ref_shim.available() == os.path.isdir(os.path.join(ref_shim.REFERENCE_ROOT, "projects", "mvsdetection"))
