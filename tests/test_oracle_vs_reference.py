"""The oracle against the UNMODIFIED reference on random scenes (CPU; only where /root/reference is mounted).

tests/golden pins the oracle on five hand-picked scenes; this sweep draws the same random family the GPU sweep uses
(tests/test_gpu_random.py: random grids, voxel sizes, float origins, intrinsics, cameras outside the grid or looking
away, three TSDF families, N in {40, 97, 300}, three thresholds) and runs the reference's own functions through
oracle/ref_shim.py next to the C oracle: Stage A indices / masks / sums / means bit-exact, ray parameters bit-exact,
NeuS rows with identical kept sets, bit-exact positions and features, weights within 1e-5.  Together with the GPU
sweep (kernels == oracle on this family) it closes the chain reference == oracle == kernels."""
import os
import types

import numpy as np
import pytest
import torch

import oracle
import ref_shim
from conftest import assert_rel, oracle_fusion, random_fusion_case, random_scene

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
SEEDS = int(os.environ.get("CNRMA_REFERENCE_SEEDS", "40"))


def _assert_differences_are_banded(ref_rows, keep, places, wraw, thr, view):
    """Walks the oracle's samples in flat (v, u, step) order -- the order of the reference's rows -- and pairs them with
    the reference's rows by bit-identical position.  A sample only the oracle keeps, or only the reference keeps, must
    have |w - thr| <= 1e-5 * thr; every reference row must be paired."""
    band = np.abs(wraw - thr) <= 1e-5 * thr
    ref_pos = np.ascontiguousarray(ref_rows[:, :3]).view(np.uint32)
    pos = np.ascontiguousarray(places).view(np.uint32)
    j = 0
    only_oracle = only_ref = 0
    for i in np.nonzero(keep | band)[0]:
        match = j < ref_pos.shape[0] and np.array_equal(pos[i], ref_pos[j])
        if keep[i] and match:
            j += 1
        elif keep[i]:
            assert band[i], f"view {view}: sample {i} kept by the oracle only, weight {wraw[i]} outside the band of {thr}"
            only_oracle += 1
        elif match:                      # band[i] holds here: the reference keeps a banded sample the oracle drops
            j += 1
            only_ref += 1
    assert j == ref_pos.shape[0], f"view {view}: {ref_pos.shape[0] - j} reference rows have no counterpart among the oracle's samples"
    assert only_oracle + only_ref > 0


def test_band_check_catches_differences_outside_the_band():
    """The checker itself: a row dropped or added outside the threshold band must fail, inside it must pass."""
    thr = 0.05
    places = np.arange(30, dtype=np.float32).reshape(10, 3)
    wraw = np.array([0.2, 0.01, 0.05 * (1 + 5e-6), 0.3, 0.05 * (1 - 5e-6), 0.0, 0.4, 0.02, 0.5, 0.6], np.float32)
    keep = wraw >= thr
    rows = lambda idx: np.concatenate([places[idx], wraw[idx, None]], 1)
    # the reference drops banded sample 2 and keeps banded sample 4: fine
    _assert_differences_are_banded(rows([0, 3, 4, 6, 8, 9]), keep, places, wraw, thr, 0)
    with pytest.raises(AssertionError):          # the reference lacks un-banded sample 3
        _assert_differences_are_banded(rows([0, 2, 6, 8, 9]), keep, places, wraw, thr, 0)
    with pytest.raises(AssertionError):          # the reference has un-banded sample 1
        _assert_differences_are_banded(rows([0, 1, 2, 3, 6, 8, 9]), keep, places, wraw, thr, 0)


@pytest.mark.parametrize("seed", range(SEEDS))
def test_oracle_equals_reference_on_random_scene(seed):
    rm = ref_shim.load_reference()
    s = random_scene(np.random.default_rng(5000 + seed))
    V, C, H, W = s["feats"].shape
    dim, vs, origin, stride, thr, N = s["dim"], s["vs"], s["origin"], s["stride"], s["thr"], s["N"]
    me = ref_shim.make_self(dim, vs, torch.from_numpy(origin).view(1, 3), stride=stride, neus_threshold=thr)
    feats = torch.from_numpy(s["feats"]).unsqueeze(1)
    projs = torch.from_numpy(s["projs"]).unsqueeze(1)
    tsdf = torch.from_numpy(s["tsdf"])[None, None]
    with torch.no_grad():
        # Stage A: per-view backproject (rm.py:21-69) and the running sum / mean (rm.py:220-257)
        for v in range(V):
            ps = projs[v].clone()
            ps[:, :2] = ps[:, :2] / stride
            vol_v, valid_v = rm.backproject(dim, vs, me.origin, ps, feats[v])
            opx, opy, ovalid = oracle.project(dim, vs, origin, oracle.scale_projection(s["projs"][v], stride), H, W)
            assert np.array_equal(valid_v[0, 0].numpy(), ovalid.reshape(dim))
            me.aggregate_2d_features(projs[v], feats[v])
        count_ref = me.valid.clone()
        me.clear_3d_features()
        ovol, ocnt = oracle.aggregate_views(s["projs"], s["feats"], dim, vs, origin, stride, mean=True)
        assert np.array_equal(count_ref[0, 0].numpy(), ocnt)
        assert np.array_equal(me.volume[0].numpy().view(np.uint32), ovol.view(np.uint32))
        # rays (rm.py:71-111)
        ps0 = projs[0].clone()
        ps0[:, :2] = ps0[:, :2] / stride
        o_ref, d_ref = rm.get_ray_parameter(ps0, feats[0])
        oo, od = oracle.rays(oracle.invert_projection(oracle.scale_projection(s["projs"][0], stride)), H, W)
        finite = np.isfinite(o_ref[0].numpy()).all() and np.isfinite(d_ref[0].numpy()).all()
        if finite:                      # a singular projection gives NaN / inf rays on both sides; bits of NaNs may differ
            assert np.array_equal(o_ref[0].numpy().view(np.uint32), oo.view(np.uint32))
            assert np.array_equal(d_ref[0].numpy().view(np.uint32), od.view(np.uint32))
        # NeuS rows per view (rm.py:687-807)
        for v in range(V):
            ps = projs[v].clone()
            ps[:, :2] = ps[:, :2] / stride
            try:
                r = me.ray_projection_neus(ps, feats[v], tsdf, grids=N, weight_threshold=thr)
            except Exception:           # rm.py:277-283: the caller swallows per-view failures ("No valid points!")
                r = None
            ro = oracle.ray_projection_neus(oracle.scale_projection(s["projs"][v], stride), s["feats"][v], s["tsdf"], dim, vs,
                                            origin, N, thr)
            if r is None or r[0] is None:
                assert ro is None or ro.shape[0] <= 1        # a single kept sample makes the reference fail (rm.py:781-783)
                continue
            ref_rows = r[0].numpy()
            if ro is None or ro.shape != ref_rows.shape:
                # kept sets may differ only in samples whose weight is within 1e-5 of the threshold (band protocol):
                # every row one side has and the other lacks is identified and its weight checked
                _w, keep, places, wraw = oracle.neus_dense(oracle.invert_projection(oracle.scale_projection(s["projs"][v], stride)),
                                                          H, W, N, dim, vs, origin, s["tsdf"], thr)
                _assert_differences_are_banded(ref_rows, keep.ravel(), places.reshape(-1, 3), wraw.ravel(), thr, v)
                continue
            assert np.array_equal(ro[:, :3].view(np.uint32), ref_rows[:, :3].view(np.uint32))
            assert np.array_equal(ro[:, 4:].view(np.uint32), ref_rows[:, 4:].view(np.uint32))
            assert_rel(ro[:, 3], ref_rows[:, 3], 1e-5, what="neus weights")


@pytest.mark.parametrize("seed", range(min(SEEDS, 16)))
def test_oracle_depth_mode_equals_reference_on_random_scene(seed):
    """The alternative weights of rm.py:809-956 (first sign change of the TSDF along the ray, triangular weights)."""
    ref_shim.load_reference()
    s = random_scene(np.random.default_rng(7000 + seed))
    dim, vs, origin, stride, N = s["dim"], s["vs"], s["origin"], s["stride"], s["N"]
    me = ref_shim.make_self(dim, vs, torch.from_numpy(origin).view(1, 3), stride=stride, ray_marching_type="depth",
                            neus_threshold=None, depth_points=2)
    feats = torch.from_numpy(s["feats"]).unsqueeze(1)
    projs = torch.from_numpy(s["projs"]).unsqueeze(1)
    tsdf = torch.from_numpy(s["tsdf"])[None, None]
    checked = 0
    with torch.no_grad():
        for k in (0, 1, 3):
            for v in range(feats.shape[0]):
                ps = projs[v].clone()
                ps[:, :2] = ps[:, :2] / stride
                try:
                    r = me.ray_projection_depth(ps, feats[v], tsdf, grids=N, select_grids=k)
                except Exception:       # rm.py:277-283 swallows per-view failures
                    r = None
                ro = oracle.ray_projection_depth(oracle.scale_projection(s["projs"][v], stride), s["feats"][v], s["tsdf"], dim,
                                                 vs, origin, N, k)
                if r is None or r[0] is None:
                    assert ro is None or ro.shape[0] <= 1
                    continue
                ref_rows = r[0].numpy()
                assert ro is not None and ro.shape == ref_rows.shape, (k, v)
                assert np.array_equal(ro.view(np.uint32), ref_rows.view(np.uint32)), (k, v)
                checked += ref_rows.shape[0]
    _DEPTH_ROWS_CHECKED.append(checked)


_DEPTH_ROWS_CHECKED = []


def test_depth_sweep_compared_something():
    """Runs after the sweep above: the random scenes must have produced depth rows to compare."""
    assert sum(_DEPTH_ROWS_CHECKED) > 1000


@pytest.mark.parametrize("seed", range(8))
def test_oracle_fusion_equals_reference_on_random_frames(seed):
    """GT TSDF fusion (data_prepare/scannet/tsdf.py:402-451): random grids, origins, frame counts and depth maps with
    holes, fused frame by frame by the unmodified reference class and by the oracle: all volumes bit-exact."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import make_golden_fusion
    ref = make_golden_fusion.load_reference_fusion()
    case = random_fusion_case(np.random.default_rng(9000 + seed))
    fus = ref.TSDFFusion(case["dim"], case["vs"], case["origin"], trunc_ratio=3, device=torch.device("cpu"), color=True, label=True)
    for i in range(case["P"].shape[0]):
        fus.integrate(torch.from_numpy(case["P"][i]), torch.from_numpy(case["depth"][i]), torch.from_numpy(case["color"][i]),
                      torch.from_numpy(case["label"][i]))
    tsdf, weight, col, lab = oracle_fusion(case)
    assert int((fus.weight_vol > 0).sum()) > 0
    assert np.array_equal(weight.view(np.uint32), fus.weight_vol.numpy().view(np.uint32))
    assert np.array_equal(tsdf.view(np.uint32), fus.tsdf_vol.numpy().view(np.uint32))
    assert np.array_equal(col.view(np.uint32), fus.color_vol.numpy().view(np.uint32))
    assert np.array_equal(lab, fus.label_vol.numpy())


@pytest.mark.parametrize("seed", range(6))
def test_oracle_head_equals_reference_on_random_volumes(seed):
    """TSDF head (models/atlas_head.py:38-52): random channel counts, extents and thresholds; values within 1e-5, the
    sparsified voxels exact, the surface mask identical outside a 1e-5 band around the threshold."""
    import importlib
    ref_shim.load_reference()
    head_mod = importlib.import_module("projects.mvsdetection.models.atlas_head")
    rng = np.random.default_rng(11000 + seed)
    torch.manual_seed(11000 + seed)
    chans = [int(c) for c in rng.integers(2, 24, size=3)]
    coarse = tuple(int(v) for v in rng.integers(2, 7, size=3))
    thr = [float(t) for t in rng.choice([0.7, 0.9, 0.99], size=3)]
    head = head_mod.AtlasTSDFHead(chans, 3, 0.04, 1.05, thr)
    xs = [float(rng.uniform(1.0, 4.0)) * torch.randn((1, c) + tuple(d * 2 ** i for d in coarse)) for i, c in enumerate(chans[::-1])]
    with torch.no_grad():
        out, _ = head(xs)
    prev = None
    for i, key in enumerate(head.keys):
        ref_t = out["scene_tsdf_" + key][0, 0].numpy()
        w = head.decoders[i].weight.detach().numpy().reshape(-1)
        # chain on the reference's previous scale so that differences do not accumulate
        t, m = oracle.tsdf_head_scale(xs[i][0].numpy(), w, prev, 1.05, thr[i - 1] if i > 0 else None)
        assert np.max(np.abs(t - ref_t)) <= 1e-5
        if i > 0:
            up = np.repeat(np.repeat(np.repeat(prev, 2, 0), 2, 1), 2, 2)
            assert np.array_equal(m, np.abs(up) < np.float32(thr[i - 1]))
            assert np.array_equal(t[~m], ref_t[~m])
        prev = ref_t
