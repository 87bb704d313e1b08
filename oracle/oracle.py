"""TEST INFRASTRUCTURE ONLY -- numpy front end of the C oracle (oracle/cnrma_oracle.c).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package.  Function names mirror the
reference's (projects/mvsdetection/models/ray_marching.py, "rm.py" below).

The one piece of arithmetic this file performs itself is the 4x4 inverse of
[P; 0 0 0 1] (rm.py:96-102), done with `torch.inverse` on CPU fp32 exactly like the
reference does, because LAPACK's LU is the one op of the path a C restatement
cannot reproduce bitwise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcnrma_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "cnrma_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.cnrma_oracle_t_one.restype = C.c_float
        _lib.cnrma_oracle_t_one.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        _lib.cnrma_oracle_neus_count.restype = C.c_int64
        _lib.cnrma_oracle_depth_rows.restype = C.c_int64
        _lib.cnrma_oracle_normalize_rows.restype = C.c_float
        _lib.cnrma_oracle_num_threads.restype = C.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return int(lib().cnrma_oracle_num_threads())


def set_threads(n):
    lib().cnrma_oracle_set_threads(C.c_int(int(n)))


def scale_projection(projection, stride):
    """rm.py:238-239: rows 0-1 of a [3,4] projection divided by the stride."""
    p = _f(projection).reshape(12)
    out = np.empty(12, np.float32)
    lib().cnrma_oracle_scale_projection(_p(p), C.c_float(stride), _p(out))
    return out.reshape(3, 4)


def project(voxel_dim, voxel_size, origin, projection_scaled, height, width):
    """rm.py:47-58 for one view -> (px int64[nvox], py int64[nvox], valid bool[nvox])."""
    nx, ny, nz = voxel_dim
    n = nx * ny * nz
    px = np.empty(n, np.int64)
    py = np.empty(n, np.int64)
    valid = np.empty(n, np.uint8)
    o = _f(origin).reshape(3)
    p = _f(projection_scaled).reshape(12)
    lib().cnrma_oracle_project(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_float(voxel_size), _p(o), _p(p),
                               C.c_int(height), C.c_int(width), _p(px), _p(py), _p(valid))
    return px, py, valid.astype(bool)


def aggregate_views(projections, features, voxel_dim, voxel_size, origin, stride, mean=True):
    """Stage A over all views (rm.py:21-69, :220-257).

    projections [V,3,4] un-scaled, features [V,C,H,W] -> (volume [C,nx,ny,nz] f32, count [nx,ny,nz] i64).
    """
    feats = _f(features)
    v, c, h, w = feats.shape
    nx, ny, nz = voxel_dim
    n = nx * ny * nz
    vol = np.empty((c, n), np.float32)
    cnt = np.empty(n, np.int64)
    p = _f(projections).reshape(v, 12)
    o = _f(origin).reshape(3)
    lib().cnrma_oracle_aggregate_views(C.c_int(v), C.c_int(c), C.c_int(h), C.c_int(w), C.c_int(nx), C.c_int(ny),
                                       C.c_int(nz), C.c_float(voxel_size), _p(o), C.c_float(stride), _p(p),
                                       _p(feats), C.c_int(1 if mean else 0), _p(vol), _p(cnt))
    return vol.reshape(c, nx, ny, nz), cnt.reshape(nx, ny, nz)


def invert_projection(projection_scaled):
    """rm.py:96-102: inverse of [P; 0 0 0 1] with torch.inverse on CPU fp32, as the reference does."""
    import torch
    p4 = torch.cat((torch.from_numpy(_f(projection_scaled).reshape(3, 4)),
                    torch.tensor([[0.0, 0.0, 0.0, 1.0]])), dim=0)
    return torch.inverse(p4).numpy().copy()


def rays(pinv, height, width):
    """rm.py:71-111 -> (o [3,H*W], d [3,H*W])."""
    o = np.empty((3, height * width), np.float32)
    d = np.empty((3, height * width), np.float32)
    lib().cnrma_oracle_rays(_p(_f(pinv).reshape(16)), C.c_int(height), C.c_int(width), _p(o), _p(d))
    return o, d


def t_one(voxel_dim, voxel_size, grids):
    x, y, z = voxel_dim
    return float(lib().cnrma_oracle_t_one(C.c_int(x), C.c_int(y), C.c_int(z), C.c_double(voxel_size), C.c_int(grids)))


def _march_args(pinv, height, width, grids, voxel_dim, voxel_size, origin, tsdf, thr):
    x, y, z = voxel_dim
    t1 = t_one(voxel_dim, voxel_size, grids)
    keepalive = (_f(pinv).reshape(16), _f(origin).reshape(3), _f(tsdf).reshape(-1))
    args = [_p(keepalive[0]), C.c_int(height), C.c_int(width), C.c_int(grids), C.c_float(t1), _p(keepalive[1]),
            C.c_float(voxel_size), C.c_int(x), C.c_int(y), C.c_int(z), _p(keepalive[2])]
    if thr is not None:
        args.append(C.c_float(thr))
    return args, keepalive


def neus_dense(pinv, height, width, grids, voxel_dim, voxel_size, origin, tsdf, thr):
    """Per-sample outputs of rm.py:729-767 for one view:
    (weights [H,W,N] masked, keep [H,W,N] bool, places [H,W,N,3], raw weights [H,W,N])."""
    n = height * width * grids
    w = np.empty(n, np.float32)
    keep = np.empty(n, np.uint8)
    places = np.empty(n * 3, np.float32)
    wraw = np.empty(n, np.float32)
    args, _ka = _march_args(pinv, height, width, grids, voxel_dim, voxel_size, origin, tsdf, thr)
    lib().cnrma_oracle_neus_dense(*args, _p(w), _p(keep), _p(places), _p(wraw))
    shp = (height, width, grids)
    return w.reshape(shp), keep.astype(bool).reshape(shp), places.reshape(shp + (3,)), wraw.reshape(shp)


def ray_projection_neus(projection_scaled, features, tsdf, voxel_dim, voxel_size, origin, grids=300,
                        weight_threshold=0.05, pinv=None):
    """rm.py:687-807 for one view, batch 1.  features [C,H,W] -> rows [M,4+C] or None when M == 0."""
    feats = _f(features)
    c, h, w = feats.shape
    if pinv is None:
        pinv = invert_projection(projection_scaled)
    per_ray = np.empty(h * w, np.int32)
    args, _ka = _march_args(pinv, h, w, grids, voxel_dim, voxel_size, origin, tsdf, weight_threshold)
    m = int(lib().cnrma_oracle_neus_count(*args, _p(per_ray)))
    if m == 0:
        return None
    rows = np.empty((m, 4 + c), np.float32)
    lib().cnrma_oracle_neus_rows(*args, C.c_int(c), _p(feats), _p(per_ray), _p(rows))
    return rows


def ray_projection_depth(projection_scaled, features, tsdf, voxel_dim, voxel_size, origin, grids=300,
                         select_grids=1, pinv=None):
    """rm.py:809-956 for one view, batch 1 -> rows [M,4+C] or None."""
    feats = _f(features)
    c, h, w = feats.shape
    if pinv is None:
        pinv = invert_projection(projection_scaled)
    args, _ka = _march_args(pinv, h, w, grids, voxel_dim, voxel_size, origin, tsdf, None)
    m = int(lib().cnrma_oracle_depth_rows(*args, C.c_int(select_grids), C.c_int(c), _p(feats), None))
    if m == 0:
        return None
    rows = np.empty((m, 4 + c), np.float32)
    lib().cnrma_oracle_depth_rows(*args, C.c_int(select_grids), C.c_int(c), _p(feats), _p(rows))
    return rows


def normalize_rows(rows):
    """rm.py:298-307: [M,4+C] -> [M,3+C] with features scaled by w/mean(w)."""
    rows = _f(rows)
    m, c4 = rows.shape
    out = np.empty((m, c4 - 1), np.float32)
    lib().cnrma_oracle_normalize_rows(C.c_int64(m), C.c_int(c4 - 4), _p(rows), _p(out))
    return out


def aggregate_2d_features_ray_marching(projections, features, tsdf, voxel_dim, voxel_size, origin, stride,
                                       grids=300, ray_marching_type="neus", neus_threshold=0.05,
                                       depth_points=None, normalize=True, pinvs=None):
    """rm.py:260-307 for batch 1: loop views in order, concatenate rows, normalise.

    projections [V,3,4] un-scaled; features [V,C,H,W]; tsdf [nx,ny,nz].
    Returns [M,3+C] (or the un-normalised [M,4+C] rows when normalize=False); None if no view kept anything.
    """
    chunks = []
    for v in range(len(projections)):
        ps = scale_projection(projections[v], stride)
        pinv = None if pinvs is None else pinvs[v]
        if ray_marching_type == "neus":
            r = ray_projection_neus(ps, features[v], tsdf, voxel_dim, voxel_size, origin, grids, neus_threshold, pinv)
        else:
            r = ray_projection_depth(ps, features[v], tsdf, voxel_dim, voxel_size, origin, grids, depth_points, pinv)
        if r is not None:
            chunks.append(r)
    if not chunks:
        return None
    rows = np.concatenate(chunks, axis=0)
    return normalize_rows(rows) if normalize else rows


def dense_rma(rows, voxel_dim, voxel_size, origin):
    """Derived operator: scatter un-normalised rows [M,4+C] into (wsum [C,nx,ny,nz], wtot [nx,ny,nz])."""
    rows = _f(rows)
    m, c4 = rows.shape
    c = c4 - 4
    x, y, z = voxel_dim
    wsum = np.zeros((c, x * y * z), np.float64)
    wtot = np.zeros(x * y * z, np.float64)
    o = _f(origin).reshape(3)
    lib().cnrma_oracle_scatter_rows(C.c_int64(m), C.c_int(c), _p(rows), _p(o), C.c_float(voxel_size), C.c_int(x),
                                    C.c_int(y), C.c_int(z), _p(wsum), _p(wtot))
    return wsum.astype(np.float32).reshape(c, x, y, z), wtot.astype(np.float32).reshape(x, y, z)


def tsdf_integrate(voxel_dim, voxel_size, origin, projection, depth, trunc_margin, tsdf, weight, color_img=None,
                   color=None, label_img=None, label=None):
    """One frame of TSDFFusion.integrate (data_prepare/scannet/tsdf.py:402-451), in place on the numpy volumes
    tsdf / weight [nvox] f32 (and color [3,nvox] f32, label [nvox] i64 when given)."""
    nx, ny, nz = voxel_dim
    h, w = depth.shape
    keep = (_f(origin).reshape(3), _f(projection).reshape(12), _f(depth),
            None if color_img is None else _f(color_img),
            None if label_img is None else np.ascontiguousarray(label_img, dtype=np.int64))
    assert tsdf.dtype == np.float32 and weight.dtype == np.float32 and tsdf.flags.c_contiguous
    lib().cnrma_oracle_tsdf_integrate(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_float(voxel_size), _p(keep[0]),
                                      _p(keep[1]), C.c_int(h), C.c_int(w), _p(keep[2]),
                                      None if keep[3] is None else _p(keep[3]),
                                      None if keep[4] is None else _p(keep[4]), C.c_float(trunc_margin),
                                      _p(tsdf), _p(weight), None if color is None else _p(color),
                                      None if label is None else _p(label))


def tsdf_head_scale(x, weight, prev, label_smoothing, sparse_threshold):
    """One decoder of AtlasTSDFHead.forward (models/atlas_head.py:38-52): x [C,nx,ny,nz] f32, weight [C],
    prev [nx/2,ny/2,nz/2] or None -> (tsdf [nx,ny,nz] f32, mask [nx,ny,nz] bool)."""
    x = _f(x)
    c, nx, ny, nz = x.shape
    w = _f(weight).reshape(c)
    p = None if prev is None else _f(prev)
    if p is not None:
        assert p.shape == (nx // 2, ny // 2, nz // 2) and nx % 2 == 0 and ny % 2 == 0 and nz % 2 == 0
    tsdf = np.empty((nx, ny, nz), np.float32)
    mask = np.empty((nx, ny, nz), np.uint8)
    lib().cnrma_oracle_tsdf_head_scale(C.c_int(c), C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(x), _p(w),
                                       None if p is None else _p(p), C.c_float(label_smoothing),
                                       C.c_float(0.0 if sparse_threshold is None else sparse_threshold), _p(tsdf), _p(mask))
    return tsdf, mask.astype(bool)


def tsdf_head(xs, weights, label_smoothing, sparse_threshold):
    """All scales, coarse to fine (atlas_head.py:38-52): returns ([tsdf_i], [mask_i for i>0])."""
    out, masks, prev = [], [], None
    for i, (x, w) in enumerate(zip(xs, weights)):
        t, m = tsdf_head_scale(x, w, prev, label_smoothing, sparse_threshold[i - 1] if i > 0 else None)
        out.append(t)
        if i > 0:
            masks.append(m)
        prev = t
    return out, masks


def quantize_unique_first(coords, feats, voxel_size):
    """The hand-off's quantisation (rm.py:330-332 + MinkowskiEngine 0.5.4, which is NOT in the reference tree: its
    published behaviour is restated here and parity with ME itself is unpinned).  `ME.utils.batch_sparse_collate`
    stores `coords / voxel_size_fcaf3d` -- a float32 tensor divided by a Python float, i.e. IEEE float32 division on CPU --
    into an int32 tensor (truncation toward zero), and `ME.SparseTensor` keeps one row per distinct coordinate.  Which
    duplicate survives is unspecified there; the contract here is the first in row order, survivors in row order.
    Returns (cells int32 [K,3], feats [K,C], coords [K,3])."""
    c = _f(coords)
    q = np.trunc(c / np.float32(voxel_size)).astype(np.int32)
    _, first = np.unique(q, axis=0, return_index=True)
    first = np.sort(first)
    return q[first], _f(feats)[first], c[first]
