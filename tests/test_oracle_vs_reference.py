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
from conftest import assert_rel, random_scene

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
SEEDS = int(os.environ.get("CNRMA_REFERENCE_SEEDS", "40"))


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
                # kept sets may differ only for weights within 1e-5 of the threshold (band protocol)
                dense_w, _keep = oracle.neus_dense(oracle.invert_projection(oracle.scale_projection(s["projs"][v], stride)), H, W, N,
                                                   dim, vs, origin, s["tsdf"], thr)
                band = np.abs(dense_w - thr) <= 1e-5 * thr
                assert band.any(), f"view {v}: kept sets differ outside the threshold band"
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
