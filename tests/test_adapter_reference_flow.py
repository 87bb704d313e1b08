"""Drop-in check of the mmdet adapter (cn-rma_b200/module.py `make_detector_class`): the
reference's OWN `forward_train` / `forward_test` (projects/mvsdetection/models/ray_marching.py:409-521, imported
unmodified through oracle/ref_shim.py) are run once on the reference class and once on the adapter class built from
it, with the same stub networks, and must produce the same losses.

There is no GPU where the reference tree is mounted, so the three functional entry points the adapter methods call are
replaced by oracle-backed CPU stand-ins for the duration of the test (tests may use the oracle; the product never
does).  What this pins is the wiring: method names, argument order and meaning, the state attributes the reference's
control flow reads (`self.volume`, `self.valid`, `self.points_detection`), deferred Stage A, stride handling.
Skipped where /root/reference is absent (the GPU box)."""
import sys
import types

import numpy as np
import pytest
import torch

import oracle
import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")

C_FEAT, STRIDE, VOXEL_DIM, VOXEL_SIZE = 8, 4, (16, 16, 8), 0.4


def _standins(monkeypatch):
    import cnrma_b200.functional as F

    def aggregate_views(projections, features, voxel_dim, voxel_size, origin, stride, mean=True, out=None, **_kw):
        assert out is None
        feats = torch.stack(list(features), 0) if not isinstance(features, torch.Tensor) else features
        vols, cnts = [], []
        for b in range(feats.shape[1]):
            v, c = oracle.aggregate_views(projections[:, b].numpy(), feats[:, b].detach().numpy(), voxel_dim, voxel_size,
                                          np.asarray(origin, np.float32).reshape(3), stride, mean=mean)
            vols.append(torch.from_numpy(v))
            cnts.append(torch.from_numpy(c))
        vol, cnt = torch.stack(vols, 0), torch.stack(cnts, 0).unsqueeze(1).to(torch.int32)
        return vol, cnt, cnt > 0

    def rma_points(projections, features, tsdf, voxel_dim, voxel_size, origin, stride, grids=300, mode="neus",
                   threshold=None, depth_points=None, normalize=True, **_kw):
        feats = torch.stack(list(features), 0) if not isinstance(features, torch.Tensor) else features
        out = []
        for b in range(feats.shape[1]):
            rows = oracle.aggregate_2d_features_ray_marching(
                projections[:, b].numpy(), feats[:, b].detach().numpy(), tsdf[b, 0].detach().numpy(), voxel_dim, voxel_size,
                np.asarray(origin, np.float32).reshape(3), stride, grids, mode, threshold, depth_points, normalize)
            out.append(torch.from_numpy(rows) if rows is not None else torch.zeros((0, 3 + feats.shape[2])))
        return out

    def switch_pointcloud(points, offsets, max_points=None, masks=None, rng=None):
        coords, feats = [], []
        for b, pts in enumerate(points):
            mask = F.sample_points(pts.shape[0], max_points, rng) if max_points is not None else np.ones(pts.shape[0], bool)
            sel = pts[torch.from_numpy(mask)]
            coords.append(sel[:, :3] + torch.as_tensor(offsets[b], dtype=torch.float32).view(1, 3))
            feats.append(sel[:, 3:])
        return coords, feats

    monkeypatch.setattr(F, "aggregate_views", aggregate_views)
    monkeypatch.setattr(F, "rma_points", rma_points)
    monkeypatch.setattr(F, "switch_pointcloud", switch_pointcloud)
    if not hasattr(np, "bool"):                       # fcaf3d_transforms.py:288 predates numpy 2
        monkeypatch.setattr(np, "bool", bool, raising=False)


def _build(cls, tmp_path, head_mod, max_points):
    """The reference constructor itself (rm.py:116-197) with the shim's mock builders, then small real stub networks."""
    torch.manual_seed(3)
    m = cls(pixel_mean=[0.0, 0.0, 0.0], pixel_std=[1.0, 1.0, 1.0], voxel_size=VOXEL_SIZE, n_scales=3,
            voxel_dim_train=VOXEL_DIM, voxel_dim_test=VOXEL_DIM, origin=[0, 0, 0], backbone2d_stride=STRIDE,
            backbone2d={}, feature_2d={}, backbone_3d={}, tsdf_head={}, detection_backbone={}, detection_head={},
            feature_transform=None, save_path=str(tmp_path), max_points=max_points, neus_threshold=0.05)
    conv = torch.nn.Conv2d(3, C_FEAT, 3, padding=1)
    m.fpn = lambda img: torch.nn.functional.avg_pool2d(conv(img), STRIDE)
    m.feature_2d = lambda x: x
    pool = torch.nn.functional.avg_pool3d
    m.backbone3d = lambda vol: [pool(vol, 4).repeat(1, 2, 1, 1, 1), pool(vol, 2), vol[:, :4]]   # coarse -> fine
    m.tsdf_head = head_mod.AtlasTSDFHead([4, C_FEAT, 2 * C_FEAT], 3, 0.04, 1.05, [0.99, 0.99, 0.99])   # keys 016 / 008 / 004 (rm.py:440 hard-codes scene_tsdf_004)
    for d in m.tsdf_head.decoders:
        torch.nn.init.normal_(d.weight, std=2.0)

    def fcaf3d_detection(self, inputs, points, test=False):           # rm.py:322-336 without MinkowskiEngine
        coords, feats, _boxes = self.switch_pointcloud(points, inputs['gt_bboxes_3d'], inputs['offset'], test)
        return {"det_loss": sum((c.abs().mean() + f.abs().mean()) for c, f in zip(coords, feats)),
                "det_rows": torch.tensor(float(sum(c.shape[0] for c in coords)))}

    m.fcaf3d_detection = types.MethodType(fcaf3d_detection, m)
    m.post_process = lambda *a, **k: []
    return m


def _inputs():
    import cnrma_b200
    rng = np.random.default_rng(7)
    views, h, w = 5, 24, 32
    extent = tuple(d * VOXEL_SIZE for d in VOXEL_DIM)
    proj = cnrma_b200.synthetic.ring_cameras(views, h, w, STRIDE, extent, rng)
    imgs = torch.from_numpy(rng.normal(size=(1, views, 3, h * STRIDE, w * STRIDE)).astype(np.float32))
    return {"scene": ["scene0000_00"], "imgs": imgs, "projection": torch.from_numpy(proj).unsqueeze(0),
            "tsdf_list": None, "gt_bboxes_3d": [torch.zeros(1, 7)], "gt_labels_3d": [torch.zeros(1)],
            "offset": [torch.tensor([0.5, -0.25, 1.0])]}


@pytest.mark.parametrize("max_points", [None, 200])
@pytest.mark.parametrize("flow", ["forward_train", "forward_test"])
def test_reference_control_flow_on_adapter_class(monkeypatch, tmp_path, flow, max_points):
    import importlib
    import cnrma_b200
    rm = ref_shim.load_reference()
    head_mod = importlib.import_module("projects.mvsdetection.models.atlas_head")
    _standins(monkeypatch)
    Adapter = cnrma_b200.make_detector_class(rm.RayMarching)
    assert Adapter.forward_train is rm.RayMarching.forward_train and Adapter.forward_test is rm.RayMarching.forward_test
    inputs = _inputs()
    results = []
    for cls in (rm.RayMarching, Adapter):
        m = _build(cls, tmp_path, head_mod, max_points)
        captured = {}
        orig = m.fcaf3d_detection

        def spy(inp, points, test=False, _orig=orig, _cap=captured):
            out = _orig(inp, points, test)
            _cap.update(out)
            return out

        m.fcaf3d_detection = spy
        np.random.seed(11)
        with torch.no_grad():
            out = getattr(m, flow)(inputs)
        vol, valid = m.volume, m.valid
        results.append((captured, vol, valid, m.points_detection, out))
    (ref_l, ref_vol, ref_valid, ref_pts, _), (my_l, my_vol, my_valid, my_pts, _) = results
    assert torch.equal(ref_valid, my_valid) and ref_valid.dtype == my_valid.dtype == torch.bool
    assert torch.equal(ref_vol, my_vol)                                   # Stage A: bit-exact (oracle == reference)
    assert len(ref_pts) == len(my_pts) == 1 and ref_pts[0].shape == my_pts[0].shape and ref_pts[0].shape[0] > 50
    assert torch.equal(ref_pts[0][:, :3], my_pts[0][:, :3])
    assert float((ref_pts[0][:, 3:] - my_pts[0][:, 3:]).abs().max()) <= 1e-5 * float(ref_pts[0][:, 3:].abs().max())
    assert float(ref_l["det_rows"]) == float(my_l["det_rows"]) == (min(max_points, ref_pts[0].shape[0]) if max_points
                                                                   else ref_pts[0].shape[0])
    assert abs(float(ref_l["det_loss"]) - float(my_l["det_loss"])) <= 1e-5 * abs(float(ref_l["det_loss"]))


@pytest.mark.parametrize("flow", ["forward_train", "forward_test"])
def test_reference_atlas_flow_on_adapter_class(monkeypatch, tmp_path, flow):
    """The reconstruction-only model (models/atlas.py): its own forward_train / forward_test on `make_atlas_class(Atlas)`
    against the reference class -- same accumulated volume, same TSDF outputs."""
    import importlib
    import cnrma_b200
    ref_shim.load_reference()
    at = importlib.import_module("projects.mvsdetection.models.atlas")
    head_mod = importlib.import_module("projects.mvsdetection.models.atlas_head")
    _standins(monkeypatch)
    Adapter = cnrma_b200.make_atlas_class(at.Atlas)
    assert Adapter.forward_train is at.Atlas.forward_train and Adapter.forward_test is at.Atlas.forward_test
    inputs = _inputs()
    outs = []
    for cls in (at.Atlas, Adapter):
        torch.manual_seed(3)
        m = cls(pixel_mean=[0.0, 0.0, 0.0], pixel_std=[1.0, 1.0, 1.0], voxel_size=VOXEL_SIZE, n_scales=3,
                voxel_dim_train=VOXEL_DIM, voxel_dim_test=VOXEL_DIM, origin=[0, 0, 0], backbone2d_stride=STRIDE,
                backbone2d={}, feature_2d={}, backbone_3d={}, tsdf_head={}, save_path=str(tmp_path))
        conv = torch.nn.Conv2d(3, C_FEAT, 3, padding=1)
        m.fpn = lambda img, conv=conv: torch.nn.functional.avg_pool2d(conv(img), STRIDE)
        m.feature_2d = lambda x: x
        pool = torch.nn.functional.avg_pool3d
        m.backbone3d = lambda vol: [pool(vol, 4).repeat(1, 2, 1, 1, 1), pool(vol, 2), vol[:, :4]]
        m.tsdf_head = head_mod.AtlasTSDFHead([4, C_FEAT, 2 * C_FEAT], 3, 0.04, 1.05, [0.99, 0.99, 0.99])
        for d in m.tsdf_head.decoders:
            torch.nn.init.normal_(d.weight, std=2.0)
        captured = {}
        m.tsdf_head.register_forward_hook(lambda _mod, _inp, out, _cap=captured: _cap.update(out[0]))
        m.post_process = lambda *a, **k: []
        with torch.no_grad():
            getattr(m, flow)(inputs)
        outs.append(captured)
    ref_o, my_o = outs
    assert set(ref_o) == set(my_o) == {"scene_tsdf_016", "scene_tsdf_008", "scene_tsdf_004"}
    for k in ref_o:
        assert torch.equal(ref_o[k], my_o[k]), k
