"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference's aggregation path on CPU for bench.py's CPU legs.

One pass = what RayMarching.forward_test does for the path (rm.py = projects/mvsdetection/models/ray_marching.py):
    Stage A   initialize_volume, V x aggregate_2d_features, clear_3d_features          rm.py:200-257, :471-483
    Stage B   aggregate_2d_features_ray_marching (NeuS, N = grids)                      rm.py:260-307, :490
through oracle/ref_shim.py (the reference's own functions, imported unmodified from /root/reference or the staged
copy oracle/_ref).  The reference hard-codes `grids=300` at its call site (rm.py:279); configurations that march a
different length (cfg 5: 256) are run through a per-view loop that passes `grids` -- the same statements as rm.py:274-307.
"""
import os
import time

import ref_shim


def available():
    return ref_shim.available()


def host_threads():
    """All the host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the affinity mask is the truth)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def reference_pass(sc, feats_nchw, view_ids, threshold, threads=None):
    """One pass over the views `view_ids` of scene `sc` (numpy inputs; `feats_nchw` [len(view_ids),C,H,W] holds the
    feature maps of exactly those views).  Returns (seconds_a, seconds_b, rows)."""
    import numpy as np
    import torch
    torch.set_num_threads(int(threads or host_threads()))
    me = ref_shim.make_self(tuple(int(v) for v in sc.voxel_dim), float(sc.voxel_size),
                            torch.from_numpy(np.asarray(sc.origin, np.float32)).view(1, 3), stride=int(sc.stride),
                            neus_threshold=float(threshold))
    ids = list(view_ids)
    if feats_nchw.shape[0] != len(ids):
        raise ValueError("feats_nchw must hold one map per sampled view")
    feats = torch.from_numpy(feats_nchw).unsqueeze(1)                                                        # [V,1,C,H,W]
    projs = torch.from_numpy(sc.projections[ids]).unsqueeze(1)                                               # [V,1,3,4]
    tsdf = torch.from_numpy(sc.tsdf)[None, None]
    with torch.no_grad():
        t0 = time.perf_counter()
        me.initialize_volume()
        for v in range(len(ids)):
            me.aggregate_2d_features(projs[v], feats[v])
        me.clear_3d_features()
        t1 = time.perf_counter()
        if int(sc.grids) == 300:
            me.aggregate_2d_features_ray_marching(projs, feats, tsdf)
            pts = me.points_detection[0]
        else:
            pts = _rma_with_grids(me, projs, feats, tsdf, int(sc.grids), float(threshold))
        t2 = time.perf_counter()
    return t1 - t0, t2 - t1, (0 if pts is None else int(pts.shape[0]))


def _rma_with_grids(me, projections, features, tsdf, grids, threshold):
    """rm.py:260-307 with the march length passed through (the reference's loop fixes it at its default)."""
    import torch
    rows = None
    stride = me.backbone2d_stride
    for v in range(projections.shape[0]):
        p = projections[v].clone()
        p[:, :2, :] = p[:, :2, :] / stride
        try:
            r = me.ray_projection_neus(p, features[v], tsdf, grids=grids, weight_threshold=threshold)
        except Exception:
            r = None
        if r is None:
            continue
        rows = r[0] if rows is None else torch.concat((rows, r[0]), dim=0)
    if rows is None:
        return None
    w = rows[:, 3:4]
    w = w / torch.mean(w)
    return torch.concat((rows[:, 0:3], rows[:, 4:] * w), dim=1)
