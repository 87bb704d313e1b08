"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    python oracle/make_golden.py

Each fixture stores the inputs and what the reference's own code returns for them
(torch 2.11 CPU fp32, via oracle/ref_shim.py):

  Stage A  rm.py:21-69 backproject per view -> px, py (where valid), valid; rm.py:220-244 running sum and
           count; rm.py:247-257 mean volume and bool valid.
  Stage B  rm.py:687-807 ray_projection_neus per view -> rows [M_v, 4+C] (or "None" marker);
           rm.py:260-307 aggregate_2d_features_ray_marching -> points [M, 3+C];
           rm.py:809-956 ray_projection_depth for depth_points 0..2 through the same aggregate call.
           Also the intermediates rm.py:71-111 (o, d) of view 0.
  Hand-off rm.py:339-407 switch_pointcloud (+ sample_points mask, numpy seed 2024) on the NeuS points.
  Grads    d(sum(volume * G)) / d features and d(sum(points * G')) / d features from the reference's own autograd.

The fixtures are the pin for oracle/cnrma_oracle.c (tests/test_oracle_golden.py) and a second
check for the CUDA path (tests/test_gpu_golden.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "cn-rma_b200"))

import ref_shim  # noqa: E402
import synthetic  # noqa: E402


def _cases():
    base = dict(stride=4)
    tiny = synthetic.make_scene("tiny", seed=1)
    yield "tiny_room", tiny, dict(base, thr=0.05, origin=[0, 0, 0])

    tiny2 = synthetic.make_scene("tiny", seed=2)
    yield "tiny_float_origin", tiny2, dict(base, thr=0.05, origin=[-0.3, 0.2, -0.1])

    small = synthetic.make_scene(dict(views=4, channels=8, height=24, width=32, voxel_dim=(16, 16, 8),
                                      voxel_size=0.3, tsdf="random", grids=64, dtype="f32"), seed=3)
    yield "small_random", small, dict(base, thr=0.05, origin=[0, 0, 0])

    room = synthetic.make_scene(dict(views=4, channels=8, height=24, width=32, voxel_dim=(16, 12, 8),
                                     voxel_size=0.3, tsdf="room", grids=300, dtype="f32"), seed=4)
    yield "small_room_thr02", room, dict(base, thr=0.2, origin=[0, 0, 0])

    # one camera in the middle of the grid (half the voxels behind it), one far outside looking away
    # (no voxel in its frustum; its rays never enter the grid so Stage B returns None for it)
    edge = synthetic.make_scene("tiny", seed=5)
    k = np.array([[57.6, 0, 32.0], [0, 57.6, 24.0], [0, 0, 1.0]])
    pose_mid = np.eye(4)
    pose_mid[:3, :3] = np.array([[0, 0, 1.0], [-1.0, 0, 0], [0, -1.0, 0]])   # x right=-y_w, y down=-z_w, fwd=+x_w
    pose_mid[:3, 3] = [1.6, 1.6, 0.8]
    pose_out = pose_mid.copy()
    pose_out[:3, 3] = [30.0, 1.6, 0.8]
    edge.projections[1] = (k @ np.linalg.inv(pose_mid)[:3]).astype(np.float32)
    edge.projections[2] = (k @ np.linalg.inv(pose_out)[:3]).astype(np.float32)
    yield "edge_behind_and_empty", edge, dict(base, thr=0.05, origin=[0, 0, 0])


def run_case(name, sc, opt):
    rm = ref_shim.load_reference()
    origin = opt["origin"]
    stride = opt["stride"]
    thr = opt["thr"]
    V, C, H, W = sc.features.shape
    nx, ny, nz = sc.voxel_dim
    feats = torch.from_numpy(sc.features).unsqueeze(1)
    projs = torch.from_numpy(sc.projections).unsqueeze(1)
    tsdf = torch.from_numpy(sc.tsdf)[None, None]
    out = dict(projections=sc.projections, features=sc.features, tsdf=sc.tsdf,
               origin=np.asarray(origin, dtype=np.float32), origin_is_int=np.array(all(isinstance(o, int) for o in origin)),
               voxel_dim=np.array(sc.voxel_dim), voxel_size=np.float64(sc.voxel_size), stride=np.int64(stride),
               grids=np.int64(sc.grids), thr=np.float64(thr))

    # ---- Stage A
    s = ref_shim.make_self(sc.voxel_dim, sc.voxel_size, origin, stride=stride, neus_threshold=thr)
    px_all = np.zeros((V, sc.nvox), np.int32)
    py_all = np.zeros((V, sc.nvox), np.int32)
    valid_all = np.zeros((V, sc.nvox), bool)
    for v in range(V):
        p = projs[v].clone()
        p[:, :2, :] = p[:, :2, :] / stride
        # the index/mask lines of backproject, rm.py:47-58, evaluated by the reference's own helpers
        coords = rm.coordinates(sc.voxel_dim, "cpu").unsqueeze(0)
        world = coords.type_as(p) * sc.voxel_size + s.origin.unsqueeze(2)
        world = torch.cat((world, torch.ones_like(world[:, :1])), dim=1)
        cam = torch.bmm(p, world)
        px = (cam[:, 0, :] / cam[:, 2, :]).round().type(torch.long)
        py = (cam[:, 1, :] / cam[:, 2, :]).round().type(torch.long)
        valid = (px >= 0) & (py >= 0) & (px < W) & (py < H) & (cam[:, 2, :] > 0)
        vol, val = rm.backproject(sc.voxel_dim, sc.voxel_size, s.origin, p, feats[v])
        assert torch.equal(val.view(1, -1), valid)
        valid_all[v] = valid[0].numpy()
        px_all[v] = np.where(valid_all[v], px[0].numpy(), 0)
        py_all[v] = np.where(valid_all[v], py[0].numpy(), 0)
        if v == 0:
            out["backproject_v0"] = vol[0].numpy()
        s.aggregate_2d_features(projs[v], feats[v])
    out.update(px=px_all, py=py_all, valid=valid_all, vol_sum=s.volume[0].numpy().copy(),
               count=s.valid[0, 0].numpy().copy())
    s.clear_3d_features()
    out.update(vol_mean=s.volume[0].numpy().copy(), valid_any=s.valid[0, 0].numpy().copy())

    # ---- Stage B (neus)
    rows_per_view = []
    m_per_view = []
    for v in range(V):
        p = projs[v].clone()
        p[:, :2, :] = p[:, :2, :] / stride
        try:
            r = s.ray_projection_neus(p, feats[v], tsdf, grids=sc.grids, weight_threshold=thr)
        except Exception:   # the reference's caller swallows these, rm.py:277-283
            r = None
        if r is None:
            m_per_view.append(-1)
        else:
            rows_per_view.append(r[0].numpy())
            m_per_view.append(r[0].shape[0])
        if v == 0:
            o, d = rm.get_ray_parameter(p, feats[v])
            out["rays_o_v0"] = o[0].numpy()
            out["rays_d_v0"] = d[0].numpy()
            p4 = torch.cat((p[0], torch.tensor([[0.0, 0.0, 0.0, 1.0]])), dim=0)
            out["pinv_v0"] = torch.inverse(p4).numpy()
    out["neus_m_per_view"] = np.array(m_per_view)
    out["neus_rows"] = np.concatenate(rows_per_view, axis=0) if rows_per_view else np.zeros((0, 4 + C), np.float32)
    if sc.grids == 300:
        s.initialize_volume()
        s.aggregate_2d_features_ray_marching(projs, feats, tsdf)
        out["neus_points"] = s.points_detection[0].numpy()
    else:
        # grids is fixed to its default at the reference's call site (rm.py:279); for other N apply
        # rm.py:298-307 to the per-view rows by hand
        rows = torch.from_numpy(out["neus_rows"])
        w = rows[:, 3:4] / torch.mean(rows[:, 3:4])
        out["neus_points"] = torch.concat((rows[:, 0:3], rows[:, 4:] * w), dim=1).numpy()

    # ---- Stage B (depth), depth_points 0..2
    for k in (0, 1, 2):
        sd = ref_shim.make_self(sc.voxel_dim, sc.voxel_size, origin, stride=stride, ray_marching_type="depth",
                                depth_points=k)
        chunks = []
        for v in range(V):
            p = projs[v].clone()
            p[:, :2, :] = p[:, :2, :] / stride
            try:
                r = sd.ray_projection_depth(p, feats[v], tsdf, grids=sc.grids, select_grids=k)
            except Exception:
                r = None
            if r is not None:
                chunks.append(r[0])
        rows = torch.concat(chunks, dim=0)
        w = rows[:, 3:4] / torch.mean(rows[:, 3:4])
        out[f"depth{k}_rows"] = rows.numpy()
        out[f"depth{k}_points"] = torch.concat((rows[:, 0:3], rows[:, 4:] * w), dim=1).numpy()

    # ---- gradients w.r.t. the feature maps (what autograd gives the reference in forward_train, rm.py:409-451)
    gen = torch.Generator().manual_seed(1234)
    fg = torch.from_numpy(sc.features).unsqueeze(1).clone().requires_grad_(True)
    sg = ref_shim.make_self(sc.voxel_dim, sc.voxel_size, origin, stride=stride, neus_threshold=thr)
    for v in range(V):
        sg.aggregate_2d_features(projs[v], fg[v])
    sg.clear_3d_features()
    g_vol = torch.randn(sg.volume.shape, generator=gen)
    (sg.volume * g_vol).sum().backward()
    out["grad_volume"] = g_vol[0].numpy()
    out["grad_features_stage_a"] = fg.grad[:, 0].numpy().copy()
    fg.grad = None
    chunks = []
    for v in range(V):
        p = projs[v].clone()
        p[:, :2, :] = p[:, :2, :] / stride
        try:
            r = sg.ray_projection_neus(p, fg[v], tsdf, grids=sc.grids, weight_threshold=thr)
        except Exception:
            r = None
        if r is not None:
            chunks.append(r[0])
    rows_g = torch.concat(chunks, dim=0)
    w_g = rows_g[:, 3:4] / torch.mean(rows_g[:, 3:4])                       # rm.py:298-307
    pts_g = torch.concat((rows_g[:, 0:3], rows_g[:, 4:] * w_g), dim=1)
    g_pts = torch.randn(pts_g.shape, generator=gen)
    (pts_g * g_pts).sum().backward()
    out["grad_points"] = g_pts.numpy()
    out["grad_features_stage_b"] = fg.grad[:, 0].numpy().copy()

    # ---- hand-off: the reference's switch_pointcloud (rm.py:339-407) on the points above, max_points = M // 3
    sh = ref_shim.make_self(sc.voxel_dim, sc.voxel_size, origin, stride=stride, neus_threshold=thr)
    sh.max_points = max(1, out["neus_points"].shape[0] // 3)
    sh.feature_transform = None
    sh.switch_pointcloud = __import__("types").MethodType(rm.RayMarching.switch_pointcloud, sh)
    offset = torch.tensor([0.37, -1.25, 0.5])
    np.random.seed(2024)
    sel_c, sel_f, _ = sh.switch_pointcloud([torch.from_numpy(out["neus_points"])], [None], [offset], True)
    np.random.seed(2024)
    out["handoff_mask"] = rm.sample_points(torch.from_numpy(out["neus_points"])[:, 0:3], max_points=sh.max_points).numpy()
    out["handoff_offset"] = offset.numpy()
    out["handoff_max_points"] = np.int64(sh.max_points)
    out["handoff_coords"] = sel_c[0].numpy()
    out["handoff_features"] = sel_f[0].numpy()

    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: V={V} C={C} HxW={H}x{W} grid={sc.voxel_dim} N={sc.grids} valid={valid_all.mean():.3f} "
          f"M={out['neus_points'].shape[0]} m_per_view={m_per_view} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    torch.manual_seed(0)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, sc, opt in _cases():
        run_case(name, sc, opt)


if __name__ == "__main__":
    main()
