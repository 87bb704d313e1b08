"""Functional host API: the reference's hot-path functions, same names and argument order, backed by the
sm_100a kernels in libcnrma_b200.so.  ("rm.py" = projects/mvsdetection/models/ray_marching.py.)

    backproject(voxel_dim, voxel_size, origin, projection, features)            rm.py:21
    get_ray_parameter(projection, features)                                      rm.py:71
plus the fused entry points the stateful mirror (module.py) is built on:
    aggregate_views(...)   = V x aggregate_2d_features + clear_3d_features      rm.py:220-257
    rma_points(...)        = aggregate_2d_features_ray_marching                  rm.py:260-307
    ray_projection(...)    = ray_projection_neus / _depth for one view           rm.py:687 / :809
    dense_rma(...)         = derived voxel-form of the RMA lift

PyTorch is used for device memory, streams and the 4x4 LAPACK inverse the reference itself calls
(rm.py:100); every other step runs in the CUDA library.  There is no CPU or eager fallback: tensors must
live on a CUDA device and the library must be built.
"""
import collections
import ctypes as C
import threading
import types

import torch

from . import _lib
from ._lib import CnrmaError

_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _origin3(origin):
    if isinstance(origin, torch.Tensor):
        return [float(v) for v in origin.detach().reshape(-1).to("cpu", torch.float32)]
    return [float(v) for v in origin]


def _as_view_list(features):
    """[V,B,C,H,W] tensor or sequence of V [B,C,H,W] tensors -> list of V tensors."""
    if isinstance(features, torch.Tensor):
        if features.dim() != 5:
            raise ValueError("features must be [V,B,C,H,W] or a sequence of [B,C,H,W]")
        return [features[v] for v in range(features.shape[0])]
    views = list(features)
    if not views or any(f.dim() != 4 for f in views):
        raise ValueError("features must be [V,B,C,H,W] or a sequence of [B,C,H,W]")
    return views


class _FeatureStack:
    """Per-batch-element cnrma_features descriptors for the views of a scene: a stacked [V,B,C,H,W] tensor (the
    pointers are then computed from its strides, no per-view tensors are made) or a list of per-view [B,C,H,W] tensors.

    Keeps the tensors (and any channels-last copies) alive for as long as the descriptor is used."""

    def __init__(self, features, need_vector_layout):
        stacked = isinstance(features, torch.Tensor)
        if stacked:
            if features.dim() != 5:
                raise ValueError("features must be [V,B,C,H,W] or a sequence of [B,C,H,W]")
            f0 = features[0]
            self.V = features.shape[0]
        else:
            views = _as_view_list(features)
            f0 = views[0]
            self.V = len(views)
        if not f0.is_cuda:
            raise CnrmaError("features must be CUDA tensors: this path has no CPU implementation")
        if f0.dtype not in _DTYPES:
            raise CnrmaError(f"unsupported feature dtype {f0.dtype} (float32 or bfloat16)")
        self.device = f0.device
        self.dtype = f0.dtype
        self.B, self.C, self.H, self.W = f0.shape
        e = 8 if self.dtype == torch.bfloat16 else 4
        if need_vector_layout and self.C % e != 0:
            raise CnrmaError(f"channels must be a multiple of {e} for {self.dtype}")
        esz = f0.element_size()
        if stacked:
            self.views = None
            self.keep = features.detach()
            strides = {tuple(features.stride()[2:])}
            ptr0, vstep = features.data_ptr(), features.stride(0) * esz
            aligned = ptr0 % 16 == 0 and vstep % 16 == 0 and (features.stride(1) * esz) % 16 == 0
        else:
            for f in views:
                if f.shape != f0.shape or f.dtype != f0.dtype or f.device != f0.device:
                    raise ValueError("all views must share shape, dtype and device")
            self.views = [f.detach() for f in views]
            self.keep = self.views
            strides = {f.stride()[1:] for f in self.views}
            aligned = all(f.data_ptr() % 16 == 0 and (f.stride(0) * esz) % 16 == 0 for f in self.views)
        sc, sy, sx = next(iter(strides))
        cl = len(strides) == 1 and sc == 1 and (not need_vector_layout or (sx % e == 0 and sy % e == 0 and aligned))
        if not cl:
            if self.views is None:
                self.views = [self.keep[v] for v in range(self.V)]
            self.views = self._to_channels_last()
            self.keep = self.views
            stacked = False
        self.converted = not cl
        if stacked:
            self.sc, self.sy, self.sx = sc, sy, sx
            self._base = [ptr0 + v * vstep for v in range(self.V)]
            self._bstride = [features.stride(1) * esz] * self.V
        else:
            self.sc, self.sy, self.sx = self.views[0].stride()[1:]
            self._base = [f.data_ptr() for f in self.views]
            self._bstride = [f.stride(0) * esz for f in self.views]

    def _to_channels_last(self):
        lib = _lib.load()
        out = _lib.empty((self.B, self.V, self.H, self.W, self.C), dtype=self.dtype, device=self.device)
        # runs of consecutive views that share strides go out in one launch per batch element
        runs, start = [], 0
        for v in range(1, self.V + 1):
            if v == self.V or self.views[v].stride() != self.views[start].stride():
                runs.append((start, v))
                start = v
        for b in range(self.B):
            for v0, v1 in runs:
                sc, sy, sx = self.views[v0].stride()[1:]
                ptrs = (C.c_void_p * (v1 - v0))(*[self.views[v][b].data_ptr() for v in range(v0, v1)])
                desc = _lib.Features(v1 - v0, self.C, self.H, self.W, _DTYPES[self.dtype], sc, sy, sx,
                                     C.cast(ptrs, C.POINTER(C.c_void_p)))
                _lib.check(lib.cnrma_to_channels_last(C.byref(desc), C.c_void_p(out[b, v0].data_ptr()),
                                                      _stream(self.device)), "cnrma_to_channels_last")
        return [out[:, v].permute(0, 3, 1, 2) for v in range(self.V)]

    def descriptor(self, b, v0=0, nv=None):
        nv = self.V - v0 if nv is None else nv
        ptrs = (C.c_void_p * nv)(*[self._base[v] + b * self._bstride[v] for v in range(v0, v0 + nv)])
        desc = _lib.Features(nv, self.C, self.H, self.W, _DTYPES[self.dtype], self.sc, self.sy, self.sx,
                             C.cast(ptrs, C.POINTER(C.c_void_p)))
        desc._keepalive = ptrs
        return desc


def _projections_device(projections, device):
    """[V,B,3,4] (tensor or sequence of [B,3,4]) -> contiguous fp32 CUDA tensor."""
    if not isinstance(projections, torch.Tensor):
        projections = torch.stack(list(projections), dim=0)
    if projections.dim() != 4 or projections.shape[-2:] != (3, 4):
        raise ValueError("projections must be [V,B,3,4]")
    return projections.detach().to(device=device, dtype=torch.float32).contiguous()


# ----------------------------------------------------------------------------------------------------
# autograd: the forward kernels run without a graph; these Functions attach the backward kernels to their
# results (the gradient flows into the feature maps only, like in the reference: rm.py:61-64, :304, :799)
# ----------------------------------------------------------------------------------------------------

def _needs_grad(views):
    return torch.is_grad_enabled() and any(v.requires_grad for v in views)


def _autograd_inputs(features, view_list):
    """What the backward Functions differentiate with respect to: the stacked [V,B,C,H,W] tensor itself when the
    caller passed one (a single gradient tensor comes back, instead of V select-backward nodes that each
    materialise a full-size zero tensor), else the per-view tensors."""
    if isinstance(features, torch.Tensor):
        return [features], True
    return view_list, False


def _grads_out(st, buf, grads):
    if st["stacked"]:
        g = buf.permute(0, 1, 4, 2, 3)
        return (None, None, g.to(st["dtypes"][0]) if st["needs"][0] else None)
    return (None, None) + tuple(gv.to(dt) if need else None for gv, dt, need in zip(grads, st["dtypes"], st["needs"]))


def _grad_buffer(state):
    """Zeroed fp32 channels-last gradient maps [V,B,H,W,C] and their per-view [B,C,H,W] views."""
    V, B, Cc, H, W = state["shape"]
    buf = torch.zeros((V, B, H, W, Cc), dtype=torch.float32, device=state["device"])
    return buf, [buf[v].permute(0, 3, 1, 2) for v in range(V)]


def _grad_descriptor(buf, b):
    V, B, H, W, Cc = buf.shape
    ptrs = (C.c_void_p * V)(*[buf[v, b].data_ptr() for v in range(V)])
    desc = _lib.Features(V, Cc, H, W, _lib.F32, 1, W * Cc, Cc, C.cast(ptrs, C.POINTER(C.c_void_p)))
    desc._keepalive = ptrs
    return desc


def _linear_voxel_strides(g, nx, ny, nz):
    """(voxel stride, channel stride) of a [C,nx,ny,nz] tensor whose voxels are linearly addressable."""
    sc, sx, sy, sz = g.stride()
    if sy == nz * sz and sx == ny * nz * sz:
        return g, sz, sc
    g = g.contiguous()
    return g, 1, nx * ny * nz


class _AggregateBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, volume, state, *views):
        ctx.state = state
        return volume.view_as(volume)

    @staticmethod
    def backward(ctx, grad_volume):
        st = ctx.state
        lib = _lib.load()
        nx, ny, nz = st["voxel_dim"]
        buf, grads = _grad_buffer(st)
        device = st["device"]
        with torch.cuda.device(device):
            for b in range(st["shape"][1]):
                g, vsv, vsc = _linear_voxel_strides(grad_volume[b].float(), nx, ny, nz)
                desc = _grad_descriptor(buf, b)
                _lib.check(lib.cnrma_aggregate_views_backward(
                    C.byref(st["grid"]), C.byref(desc), C.c_void_p(st["P"][0, b].data_ptr()), st["shape"][1] * 12,
                    float(st["stride"]), _lib.AGG_MEAN if st["mean"] else 0, C.c_void_p(g.data_ptr()), vsv, vsc,
                    C.c_void_p(st["count"][b].data_ptr()), _stream(device)), "cnrma_aggregate_views_backward")
        return _grads_out(st, buf, grads)


class _RmaRowsBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, state, *views):
        ctx.state = state
        return rows.view_as(rows)

    @staticmethod
    def backward(ctx, grad_rows):
        st = ctx.state
        lib = _lib.load()
        m = st["march"]
        buf, grads = _grad_buffer(st)
        device = st["device"]
        g = grad_rows.float().contiguous()
        mean_t = st["mean"]
        with torch.cuda.device(device):
            desc = _grad_descriptor(buf, st["b"])
            if g.shape[0] > 0:
                _lib.check(lib.cnrma_rma_fill_backward(
                    C.byref(st["grid"]), C.byref(desc), m.grids, m.mode, m.threshold, m.depth_points,
                    C.c_void_p(m.workspace.data_ptr()), C.c_void_p(m.result.data_ptr()), 1 if st["normalize"] else 0,
                    C.c_void_p(mean_t.data_ptr()) if mean_t is not None else None, C.c_void_p(g.data_ptr()),
                    g.shape[1], _stream(device)), "cnrma_rma_fill_backward")
        return _grads_out(st, buf, grads)


# ----------------------------------------------------------------------------------------------------
# Stage A
# ----------------------------------------------------------------------------------------------------

def project_views(projections, voxel_dim, voxel_size, origin, stride, height, width, device=None):
    """Index / mask part of backproject (rm.py:47-58) for all views.

    projections [V,B,3,4] un-scaled -> (px, py int32 [V,B,nvox], valid bool [V,B,nvox]); bit-exact."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else projections.device
    P = _projections_device(projections, device)
    V, B = P.shape[:2]
    nvox = int(voxel_dim[0]) * int(voxel_dim[1]) * int(voxel_dim[2])
    grid = _lib.make_grid(voxel_dim, voxel_size, _origin3(origin))
    px = _lib.empty((B, V, nvox), dtype=torch.int32, device=device)
    py = _lib.empty_like(px)
    valid = _lib.empty((B, V, nvox), dtype=torch.bool, device=device)
    with torch.cuda.device(device):
        for b in range(B):
            _lib.check(lib.cnrma_project_views(C.byref(grid), C.c_void_p(P[0, b].data_ptr()), B * 12, V,
                                               float(stride), int(height), int(width), C.c_void_p(px[b].data_ptr()),
                                               C.c_void_p(py[b].data_ptr()), C.c_void_p(valid[b].data_ptr()),
                                               _stream(device)), "cnrma_project_views")
    return px.transpose(0, 1), py.transpose(0, 1), valid.transpose(0, 1)


def aggregate_views(projections, features, voxel_dim, voxel_size, origin, stride, mean=True, out=None,
                    accumulate=None, count_f32=False, box=None, reserve_ctas=0):
    """Fused Stage A: V x aggregate_2d_features (+ clear_3d_features when mean=True), rm.py:220-257.

    projections [V,B,3,4] un-scaled; features [V,B,C,H,W] or a sequence of [B,C,H,W].
    Returns (volume [B,C,nx,ny,nz] fp32, count [B,1,nx,ny,nz] int32, valid [B,1,nx,ny,nz] bool).
    `volume` is a channels_last_3d view of a [B,nx,ny,nz,C] buffer (logical NCDHW like the reference).
    `out=(volume, count, valid)` supplies the buffers; with `accumulate` (default when `out` is given) the
    new views are added to the un-averaged sums / counts already there (rm.py:243-244 semantics) and
    `mean` finishes them.  `count_f32` stores the counts as float32 (exact below 2**24) so sums and counts
    can share one fp32 all-reduce buffer (distributed.py).  `box=(lo, dim)` restricts the pass to the voxels
    [lo, lo + dim) of the grid: the outputs then have the box's extents, and every voxel gets the bits the full-grid
    call gives it (the unit of work of the voxel-sharded and chunk-pipelined multi-GPU modes, distributed.py);
    `reserve_ctas` leaves that many CTA slots to a kernel running beside this one."""
    lib = _lib.load()
    fs = _FeatureStack(features, need_vector_layout=True)
    device = fs.device
    P = _projections_device(projections, device)
    if P.shape[0] != fs.V or P.shape[1] != fs.B:
        raise ValueError("projections / features disagree on views or batch")
    grid = _lib.make_grid(voxel_dim, voxel_size, _origin3(origin))
    box_c = None
    if box is not None:
        lo, dim = box
        if any(int(l) < 0 or int(d) <= 0 or int(l) + int(d) > int(n) for l, d, n in zip(lo, dim, voxel_dim)):
            raise ValueError("box must lie inside the grid")
        box_c = _lib.make_box(lo, dim)
        voxel_dim = tuple(int(d) for d in dim)
    nx, ny, nz = (int(v) for v in voxel_dim)
    flags = (_lib.AGG_MEAN if mean else 0) | (_lib.AGG_COUNT_F32 if count_f32 else 0)
    if out is None:
        buf = _lib.empty((fs.B, nx, ny, nz, fs.C), dtype=torch.float32, device=device)
        count = _lib.empty((fs.B, 1, nx, ny, nz), dtype=torch.float32 if count_f32 else torch.int32, device=device)
        valid = _lib.empty((fs.B, 1, nx, ny, nz), dtype=torch.bool, device=device)
        volume = buf.permute(0, 4, 1, 2, 3)
    else:
        volume, count, valid = out
        if accumulate is None or accumulate:
            flags |= _lib.AGG_ACCUMULATE
        _check_out(volume, count, fs.B, fs.C, nx, ny, nz, count_f32)
    view_list = [features] if isinstance(features, torch.Tensor) else _as_view_list(features)
    with_grad = _needs_grad(view_list)
    if with_grad and (out is not None or count_f32 or box is not None):
        raise CnrmaError("autograd through aggregate_views needs the one-call form (no `out=` accumulation, no box)")
    with torch.cuda.device(device):
        for b in range(fs.B):
            desc = fs.descriptor(b)
            vb = volume[b]
            tail = (C.c_void_p(vb.data_ptr()), vb.stride(3), vb.stride(0), C.c_void_p(count[b].data_ptr()),
                    C.c_void_p(valid[b].data_ptr()) if valid is not None else None)
            if box_c is None:
                _lib.check(lib.cnrma_aggregate_views(C.byref(grid), C.byref(desc), C.c_void_p(P[0, b].data_ptr()),
                                                     fs.B * 12, float(stride), flags, *tail, _stream(device)),
                           "cnrma_aggregate_views")
            else:
                _lib.check(lib.cnrma_aggregate_views_box(C.byref(grid), C.byref(box_c), C.byref(desc),
                                                         C.c_void_p(P[0, b].data_ptr()), fs.B * 12, float(stride), flags,
                                                         *tail, int(reserve_ctas), _stream(device)),
                           "cnrma_aggregate_views_box")
    if with_grad:
        inputs, stacked = _autograd_inputs(features, view_list)
        state = dict(shape=(fs.V, fs.B, fs.C, fs.H, fs.W), device=device, voxel_dim=(nx, ny, nz), grid=grid, P=P,
                     stride=stride, mean=bool(mean), count=count, stacked=stacked, dtypes=[v.dtype for v in inputs],
                     needs=[v.requires_grad for v in inputs])
        volume = _AggregateBackward.apply(volume, state, *inputs)
    return volume, count, valid


def aggregate_views_bilinear(projections, features, voxel_dim, voxel_size, origin, stride, mean=True):
    """OPT-IN extra (not in the reference, which samples nearest): Stage A with bilinear sampling at the projected
    position, same validity mask / counts as aggregate_views.  Returns (volume, count, valid) in the same layouts."""
    lib = _lib.load()
    fs = _FeatureStack(features, need_vector_layout=True)
    device = fs.device
    P = _projections_device(projections, device)
    nx, ny, nz = (int(v) for v in voxel_dim)
    grid = _lib.make_grid(voxel_dim, voxel_size, _origin3(origin))
    buf = _lib.empty((fs.B, nx, ny, nz, fs.C), dtype=torch.float32, device=device)
    count = _lib.empty((fs.B, 1, nx, ny, nz), dtype=torch.int32, device=device)
    valid = _lib.empty((fs.B, 1, nx, ny, nz), dtype=torch.bool, device=device)
    with torch.cuda.device(device):
        for b in range(fs.B):
            desc = fs.descriptor(b)
            _lib.check(lib.cnrma_aggregate_views_bilinear(C.byref(grid), C.byref(desc), C.c_void_p(P[0, b].data_ptr()),
                                                          fs.B * 12, float(stride), _lib.AGG_MEAN if mean else 0,
                                                          C.c_void_p(buf[b].data_ptr()), C.c_void_p(count[b].data_ptr()),
                                                          C.c_void_p(valid[b].data_ptr()), _stream(device)),
                       "cnrma_aggregate_views_bilinear")
    return buf.permute(0, 4, 1, 2, 3), count, valid


def _check_out(volume, count, B, Cc, nx, ny, nz, count_f32):
    if tuple(volume.shape) != (B, Cc, nx, ny, nz) or volume.dtype != torch.float32 or not volume.is_cuda:
        raise ValueError("out volume has the wrong shape, dtype or device")
    # the kernel addresses voxel (x, y, z) at ((x*ny + y)*nz + z) * stride(4): the volume must be linear over the voxels
    sz = volume.stride(4)
    if (ny > 1 and volume.stride(3) != nz * sz) or (nx > 1 and volume.stride(2) != ny * nz * sz):
        raise ValueError("out volume must be linearly addressable over (x, y, z)")
    if count.dtype != (torch.float32 if count_f32 else torch.int32) or count.numel() != B * nx * ny * nz \
            or not count.is_contiguous():
        raise ValueError("out count has the wrong shape or dtype")


def finalize_views(volume, count, valid=None, count_f32=False):
    """clear_3d_features (rm.py:247-257) on sums / counts produced elsewhere -- e.g. after the all-reduce of
    view-sharded partial results: volume <- volume / count in place, 0 where count == 0."""
    lib = _lib.load()
    B, Cc, nx, ny, nz = volume.shape
    _check_out(volume, count, B, Cc, nx, ny, nz, count_f32)
    device = volume.device
    grid = _lib.make_grid((nx, ny, nz), 1.0, (0.0, 0.0, 0.0))
    flags = _lib.AGG_ACCUMULATE | _lib.AGG_MEAN | (_lib.AGG_COUNT_F32 if count_f32 else 0)
    desc = _lib.Features(0, Cc, 1, 1, _lib.F32, 1, Cc, Cc, None)
    with torch.cuda.device(device):
        for b in range(B):
            vb = volume[b]
            _lib.check(lib.cnrma_aggregate_views(C.byref(grid), C.byref(desc), None, 12, 1.0, flags,
                                                 C.c_void_p(vb.data_ptr()), vb.stride(3), vb.stride(0),
                                                 C.c_void_p(count[b].data_ptr()),
                                                 C.c_void_p(valid[b].data_ptr()) if valid is not None else None,
                                                 _stream(device)), "cnrma_aggregate_views(finalise)")
    return volume


def backproject(voxel_dim, voxel_size, origin, projection, features):
    """Drop-in for rm.py:21-69.  projection [B,3,4] ALREADY divided by the stride (as the reference's caller
    does, rm.py:238-239), features [B,C,H,W] -> (volume [B,C,nx,ny,nz], valid [B,1,nx,ny,nz] bool)."""
    volume, _count, valid = aggregate_views(projection.unsqueeze(0), features.unsqueeze(0), voxel_dim, voxel_size,
                                            origin, 1.0, mean=False)
    return volume, valid


# ----------------------------------------------------------------------------------------------------
# Stage B
# ----------------------------------------------------------------------------------------------------

_inverse_lock = threading.Lock()


def invert_projections(projections_scaled):
    """rm.py:96-102: inverse of [P; 0 0 0 1] per view with the reference's own LAPACK call (torch.inverse, CPU,
    fp32) on the host, so the ray parameters are bit-identical to the reference RUN ON THE CPU (a reference run on a GPU
    inverts with cuSOLVER / MAGMA, whose last bits differ: parity of Stage B is defined against the CPU run, as
    SURVEY.md section 8 does).  The batched call runs the same per-matrix getrf/getri as the reference's one-matrix
    calls (bit-identical, checked in the tests); it is issued single-threaded because OpenMP fan-out over fifty 4x4
    matrices costs milliseconds -- the thread count is a process-wide torch setting, so the toggle is serialised.
    Pass host tensors when the cameras are known on the host (they come from the data loader): a CUDA tensor is
    copied back first, which synchronises the stream.
    projections_scaled [V,3,4] (any device) -> [V,4,4] CPU fp32 (pinned when CUDA is available)."""
    P = projections_scaled.detach().to("cpu", torch.float32)
    V = P.shape[0]
    p4 = torch.zeros((V, 4, 4), dtype=torch.float32)
    p4[:, :3, :] = P
    p4[:, 3, 3] = 1.0
    out = torch.empty((V, 4, 4), dtype=torch.float32, pin_memory=torch.cuda.is_available())
    _inverse_lock.acquire()
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        try:
            torch.inverse(p4, out=out)
        except RuntimeError:
            # a singular camera matrix: the reference's per-view torch.inverse raises, its caller swallows the
            # exception and drops that view (rm.py:277-283).  NaN ray parameters reproduce "contributes nothing".
            for v in range(V):
                try:
                    torch.inverse(p4[v], out=out[v])
                except RuntimeError:
                    out[v].fill_(float("nan"))
    finally:
        torch.set_num_threads(threads)
        _inverse_lock.release()
    return out


def scale_projections(projections, stride):
    """rm.py:238-239 / :275-276 on the host copy: rows 0-1 divided by the backbone stride."""
    P = projections.detach().to("cpu", torch.float32).clone()
    P[..., :2, :] = P[..., :2, :] / stride
    return P


def get_ray_parameter(projection, features):
    """Drop-in for rm.py:71-111.  projection [B,3,4] (already stride-scaled), features [B,C,H,W] ->
    (o, d) each [B,3,H*W]."""
    lib = _lib.load()
    if not features.is_cuda:
        raise CnrmaError("features must be a CUDA tensor")
    device = features.device
    B, _c, H, W = features.shape
    pinv = invert_projections(projection).to(device)
    o = _lib.empty((B, 3, H * W), dtype=torch.float32, device=device)
    d = _lib.empty_like(o)
    with torch.cuda.device(device):
        _lib.check(lib.cnrma_ray_parameters(C.c_void_p(pinv.data_ptr()), B, H, W, C.c_void_p(o.data_ptr()),
                                            C.c_void_p(d.data_ptr()), _stream(device)), "cnrma_ray_parameters")
    return o, d


class _March:
    """State of one marched batch element: workspace, result block, and what fill/scatter need."""
    pass


def prepare_pinv(projections_scaled_b, device):
    """Host half of get_ray_parameter (rm.py:96-102) for one batch element: [V,3,4] scaled projections ->
    device tensor [V,4,4].  Separate so that callers can do it before queueing unrelated GPU work."""
    return invert_projections(projections_scaled_b).to(device, non_blocking=True)


def _march(fs, b, P_scaled_b, tsdf_b, grid, voxel_dim, voxel_size, grids, mode, threshold, depth_points, pinv=None):
    lib = _lib.load()
    device = fs.device
    m = _March()
    m.mode = _lib.MARCH_NEUS if mode == "neus" else _lib.MARCH_DEPTH
    m.threshold = float(threshold if threshold is not None else 0.0)
    m.depth_points = int(depth_points if depth_points is not None else 0)
    m.grids = int(grids)
    m.pinv = pinv if pinv is not None else prepare_pinv(P_scaled_b, device)
    m.t_one = lib.cnrma_t_one(C.byref(grid), float(voxel_size), m.grids)
    nbytes = C.c_size_t(0)
    _lib.check(lib.cnrma_rma_workspace_bytes(C.byref(grid), fs.V, fs.H, fs.W, m.grids, m.mode, m.threshold, m.depth_points,
                                             C.byref(nbytes)), "cnrma_rma_workspace_bytes")
    m.workspace = _lib.empty(nbytes.value, dtype=torch.uint8, device=device)
    m.result = _lib.empty(C.sizeof(_lib.RmaResult), dtype=torch.uint8, device=device)
    if tsdf_b.dtype != torch.float32 or not tsdf_b.is_contiguous() or tsdf_b.device != device:
        tsdf_b = tsdf_b.detach().to(device=device, dtype=torch.float32).contiguous()
    if tuple(tsdf_b.shape) != tuple(int(v) for v in voxel_dim):
        raise ValueError("tsdf must be [B,1,nx,ny,nz]")
    m.tsdf = tsdf_b
    _lib.check(lib.cnrma_rma_march(C.byref(grid), C.c_void_p(m.pinv.data_ptr()), fs.V, fs.H, fs.W,
                                   C.c_void_p(tsdf_b.data_ptr()), m.grids, m.t_one, m.mode, m.threshold,
                                   m.depth_points, C.c_void_p(m.workspace.data_ptr()), nbytes.value,
                                   C.c_void_p(m.result.data_ptr()), _stream(device)), "cnrma_rma_march")
    m.done = torch.cuda.Event()
    m.done.record(torch.cuda.current_stream(device))
    return m


# Host-side state of the path.  All of it is either per thread (the pinned staging slot of the M read-back) or a
# guarded hint (how many rows recent calls with the same shapes kept), so the functions of this module may be called
# from several Python threads, each on its own stream -- the same contract as the C ABI underneath.
_tls = threading.local()
_state_lock = threading.Lock()
_fill_counters = {"calls": 0, "speculative": 0, "misses": 0}
_phase_hook = None


def set_phase_hook(fn):
    """Observer for profiling: fn(label) is called right after the kernels of a phase have been queued on the current
    stream ("march" inside rma_points / rma_points_selected).  bench.py records CUDA events there; None removes it."""
    global _phase_hook
    _phase_hook = fn


def fill_stats(reset=False):
    """Counters of the speculative fill (see _march_and_fill): calls, launches made before M was known, and misses
    (the row count exceeded the guess and the fill ran a second time)."""
    with _state_lock:
        out = dict(_fill_counters)
        if reset:
            for k in _fill_counters:
                _fill_counters[k] = 0
    return out


def _read_result(m):
    """The one host sync of the path: M (and the weight sum) back to the host.  The copy runs on a side stream
    that waits only for the march / scan kernels (event recorded by _march), through a pinned staging buffer, so
    a fill kernel queued speculatively behind the march does not delay it."""
    dev = m.result.device
    slots = getattr(_tls, "pinned", None)
    if slots is None:
        slots = _tls.pinned = {}
    slot = slots.get(dev)
    if slot is None:
        slot = (torch.empty(C.sizeof(_lib.RmaResult), dtype=torch.uint8, pin_memory=True), torch.cuda.Event(),
                torch.cuda.Stream(device=dev))
        slots[dev] = slot
    host, event, side = slot
    side.wait_event(m.done)
    with torch.cuda.stream(side):
        host.copy_(m.result, non_blocking=True)
        event.record(side)
    m.result.record_stream(side)
    event.synchronize()
    res = _lib.RmaResult.from_buffer_copy(host.numpy().tobytes())
    if res.overflow:
        raise CnrmaError(f"{res.overflow} rays exceeded the per-ray record capacity")
    return res


def _fill(fs, b, m, grid, rows_host, normalize, mean_tensor=None, desc=None, capacity=None):
    """Launches the fill kernel.  With `capacity` and rows_host = -1 the launch is speculative: it happens before M
    is known on the host, into a buffer of `capacity` rows (rows beyond it are dropped by the kernel)."""
    lib = _lib.load()
    device = fs.device
    cols = fs.C + (3 if normalize else 4)
    cap = rows_host if capacity is None else capacity
    rows = _lib.empty((cap, cols), dtype=torch.float32, device=device)
    if cap == 0:
        return rows
    desc = desc if desc is not None else fs.descriptor(b)
    mean_ptr = C.c_void_p(mean_tensor.data_ptr()) if mean_tensor is not None else None
    _lib.check(lib.cnrma_rma_fill(C.byref(grid), C.c_void_p(m.pinv.data_ptr()), C.byref(desc), m.grids, m.t_one,
                                  m.mode, m.threshold, m.depth_points, C.c_void_p(m.workspace.data_ptr()),
                                  C.c_void_p(m.result.data_ptr()), rows_host, 1 if normalize else 0, mean_ptr,
                                  C.c_void_p(rows.data_ptr()), cols, cap, _stream(device)), "cnrma_rma_fill")
    return rows


# rows kept by recent calls with the same shapes: lets the fill kernel be queued before M has been read back
_rows_hint = {}


def _hint_rows(key):
    """Capacity guess for `key`: the largest row count of its last eight calls plus 6 % headroom, or None."""
    with _state_lock:
        seen = _rows_hint.get(key)
        if not seen:
            return None
        n = max(seen)
    return n + n // 16 + 1024


def _note_rows(key, n):
    with _state_lock:
        _rows_hint.setdefault(key, collections.deque(maxlen=8)).append(int(n))


def _march_and_fill(fs, b, m, grid, normalize, mean_hook, key, desc=None):
    """march result -> rows.  If earlier calls with the same shapes tell how many rows to expect, the fill is
    launched speculatively into a buffer sized for the largest of them (+ 6 %) and the read-back of M (the path's one
    host sync) overlaps with it; otherwise (first call, view-sharded mean, or a scene that keeps more rows than any
    recent one) M is read first.  fill_stats() counts how often each happens."""
    device = fs.device
    desc = desc if desc is not None else fs.descriptor(b)
    hint = _hint_rows(key)
    rows = None
    mean_t = None
    on_device = mean_hook is not None and normalize and getattr(mean_hook, "on_device", False)
    if on_device:
        # the divisor is computed from the march's result block IN DEVICE MEMORY (e.g. all-reduced over the ranks of a
        # view-sharded scene): nothing waits for the host, the fill can still be queued before M is known
        mean_t = mean_hook(m.result)
    speculative = bool(hint) and (mean_hook is None or on_device or not normalize)
    if speculative:
        rows = _fill(fs, b, m, grid, -1, normalize, mean_t, desc, capacity=hint)
    res = _read_result(m)
    n = int(res.rows)
    missed = rows is not None and n > rows.shape[0]
    if rows is None or missed:
        if mean_t is None and mean_hook and normalize:
            mean_t = mean_hook(res.weight_sum, res.rows, device)
        rows = _fill(fs, b, m, grid, n, normalize, mean_t, desc)
    _note_rows(key, n)
    with _state_lock:
        _fill_counters["calls"] += 1
        _fill_counters["speculative"] += 1 if speculative else 0
        _fill_counters["misses"] += 1 if missed else 0
    return rows[:n], res, mean_t


def _check_mode(mode, threshold, depth_points):
    if mode == "neus":
        if threshold is None:
            raise ValueError("neus ray marching needs a weight threshold (rm.py:191)")
    elif mode == "depth":
        if depth_points is None or depth_points < 0:
            raise ValueError("depth ray marching needs depth_points >= 0")
    else:
        raise ValueError("ray_marching_type must be 'neus' or 'depth'")


def rma_points(projections, features, tsdf, voxel_dim, voxel_size, origin, stride, grids=300, mode="neus",
               threshold=None, depth_points=None, normalize=True, return_stats=False, mean_hook=None):
    """aggregate_2d_features_ray_marching (rm.py:260-307) for all views at once.

    projections [V,B,3,4] un-scaled, features [V,B,C,H,W] (or sequence of [B,C,H,W]), tsdf [B,1,nx,ny,nz].
    Returns a list over the batch of [M,3+C] tensors (rows [x,y,z, feat*w/mean(w)] in (view, v, u, step)
    order); with normalize=False the un-normalised [M,4+C] rows ([x,y,z,w,feat]) the per-view function
    returns.  A batch element with no kept sample yields an empty [0, .] tensor (the reference raises).
    `mean_hook(weight_sum, rows, device) -> float32 CUDA tensor [1]` overrides the divisor of rm.py:303; a hook with the
    attribute `on_device = True` is called as `mean_hook(result_block)` instead, with the march's 32-byte result block
    (cnrma_rma_result: rows int64 at byte 0, weight_sum double at byte 8) as a uint8 CUDA tensor, and must not
    synchronise -- the fill is then queued without waiting for the host (distributed.rma_points_sharded)."""
    _check_mode(mode, threshold, depth_points)
    view_list = [features] if isinstance(features, torch.Tensor) else _as_view_list(features)
    with_grad = _needs_grad(view_list)
    fs = _FeatureStack(features, need_vector_layout=False)
    device = fs.device
    if not isinstance(projections, torch.Tensor):
        projections = torch.stack(list(projections), dim=0)
    P_scaled = scale_projections(projections, stride)            # host, [V,B,3,4]
    grid = _lib.make_grid(voxel_dim, voxel_size, _origin3(origin))
    out, stats = [], []
    with torch.cuda.device(device):
        for b in range(fs.B):
            m = _march(fs, b, P_scaled[:, b], tsdf[b, 0], grid, voxel_dim, voxel_size, grids, mode, threshold,
                       depth_points)
            # view-sharded callers replace the local mean weight by the all-reduced one (distributed.py)
            key = (device.index, b, fs.V, fs.C, fs.H, fs.W, tuple(int(v) for v in voxel_dim), int(grids), mode,
                   threshold, depth_points, bool(normalize))
            if _phase_hook is not None:
                _phase_hook("march")
            rows, res, mean_t = _march_and_fill(fs, b, m, grid, normalize, mean_hook, key)
            if with_grad:
                inputs, stacked = _autograd_inputs(features, view_list)
                state = dict(shape=(fs.V, fs.B, fs.C, fs.H, fs.W), device=device, grid=grid, march=m, b=b,
                             normalize=bool(normalize), mean=mean_t, stacked=stacked,
                             dtypes=[v.dtype for v in inputs], needs=[v.requires_grad for v in inputs])
                rows = _RmaRowsBackward.apply(rows, state, *inputs)
            out.append(rows)
            stats.append(dict(rows=int(res.rows), weight_sum=float(res.weight_sum), mean=float(res.mean)))
    return (out, stats) if return_stats else out


def ray_projection(projection, features, tsdf, voxel_dim, voxel_size, origin, grids=300, mode="neus",
                   threshold=None, depth_points=None):
    """ray_projection_neus / ray_projection_depth for ONE view (rm.py:687-807 / :809-956).
    projection [B,3,4] already stride-scaled, features [B,C,H,W] -> list over b of [M,4+C], or None if any
    batch element keeps no sample (rm.py:782-783)."""
    rows = rma_points(projection.unsqueeze(0), features.unsqueeze(0), tsdf, voxel_dim, voxel_size, origin, 1.0,
                      grids=grids, mode=mode, threshold=threshold, depth_points=depth_points, normalize=False)
    if any(r.shape[0] == 0 for r in rows):
        return None
    return rows


def rma_dense_weights(projections, height, width, tsdf, voxel_dim, voxel_size, origin, stride, grids=300,
                      threshold=0.05, device=None):
    """Parity surface for rm.py:765-767: (weights * valid_final, valid_final), each [V,H,W,N], batch 1."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else tsdf.device
    # _march only reads the geometry of the feature stack (no feature data is touched by the march)
    fs = types.SimpleNamespace(device=device, V=projections.shape[0], H=int(height), W=int(width))
    P_scaled = scale_projections(projections, stride)
    grid = _lib.make_grid(voxel_dim, voxel_size, _origin3(origin))
    with torch.cuda.device(device):
        m = _march(fs, 0, P_scaled[:, 0], tsdf[0, 0], grid, voxel_dim, voxel_size, grids, "neus", threshold, None)
        n = fs.V * fs.H * fs.W * int(grids)
        w = _lib.empty(n, dtype=torch.float32, device=device)
        keep = _lib.empty(n, dtype=torch.bool, device=device)
        _lib.check(lib.cnrma_rma_expand(fs.V, fs.H, fs.W, int(grids), float(threshold),
                                        C.c_void_p(m.workspace.data_ptr()), C.c_void_p(w.data_ptr()),
                                        C.c_void_p(keep.data_ptr()), _stream(device)), "cnrma_rma_expand")
        _read_result(m)
    shape = (fs.V, fs.H, fs.W, int(grids))
    return w.view(shape), keep.view(shape)


def dense_rma(projections, features, tsdf, voxel_dim, voxel_size, origin, stride, grids=300, mode="neus",
              threshold=None, depth_points=None, out=None):
    """Derived dense operator: per-voxel weighted feature sums and weight totals of the RMA lift.

    Returns (wsum [B,C,nx,ny,nz] fp32 (channels_last_3d view), wtot [B,1,nx,ny,nz] fp32), un-normalised so
    that view-sharded partial results add (one all-reduce, see distributed.py).  `out=(wsum, wtot)`
    accumulates into existing buffers."""
    _check_mode(mode, threshold, depth_points)
    lib = _lib.load()
    fs = _FeatureStack(features, need_vector_layout=False)
    device = fs.device
    if not isinstance(projections, torch.Tensor):
        projections = torch.stack(list(projections), dim=0)
    P_scaled = scale_projections(projections, stride)
    nx, ny, nz = (int(v) for v in voxel_dim)
    grid = _lib.make_grid(voxel_dim, voxel_size, _origin3(origin))
    if out is None:
        buf = torch.zeros((fs.B, nx, ny, nz, fs.C), dtype=torch.float32, device=device)
        wtot = torch.zeros((fs.B, 1, nx, ny, nz), dtype=torch.float32, device=device)
        wsum = buf.permute(0, 4, 1, 2, 3)
    else:
        wsum, wtot = out
        if wsum.stride(1) != 1 or not wtot.is_contiguous():
            raise ValueError("out buffers must be channels-last wsum and contiguous wtot")
    with torch.cuda.device(device):
        for b in range(fs.B):
            m = _march(fs, b, P_scaled[:, b], tsdf[b, 0], grid, voxel_dim, voxel_size, grids, mode, threshold,
                       depth_points)
            desc = fs.descriptor(b)
            _lib.check(lib.cnrma_rma_scatter(C.byref(grid), C.c_void_p(m.pinv.data_ptr()), C.byref(desc), m.grids,
                                             m.t_one, m.mode, m.threshold, m.depth_points,
                                             C.c_void_p(m.workspace.data_ptr()), C.c_void_p(wsum[b].data_ptr()),
                                             C.c_void_p(wtot[b].data_ptr()), _stream(device)), "cnrma_rma_scatter")
    return wsum, wtot


# ----------------------------------------------------------------------------------------------------
# Point-cloud hand-off (rm.py:339-407)
# ----------------------------------------------------------------------------------------------------

def sample_points(n, max_points, rng=None):
    """The sub-sampling mask of datasets/pipelines/fcaf3d_transforms.py:283-296, restated call for call so that the
    same numpy RNG state gives the same mask: a bool array [n] with max_points ones when n > max_points, all ones
    otherwise.  Host-side by design -- the reference draws it with numpy, and only the same mask gives the same rows."""
    import numpy as np
    rng = np.random if rng is None else rng
    mask = np.ones(n, dtype=bool)
    indice = np.nonzero(mask)[0]
    if n > max_points:
        choices = rng.choice(n, max_points, replace=False)
        indice = indice[choices]
    new_mask = np.zeros_like(mask)
    new_mask[indice] = 1
    return new_mask


def sample_points_device(n, max_points, seed, device):
    """On-device counterpart of sample_points: a bool CUDA tensor [n] with exactly min(n, max_points) ones, a uniform
    random subset keyed by `seed` (equivalent in distribution to the reference's numpy draw, not the same stream --
    use sample_points for parity).  Avoids the host-side permutation of n indices (~0.2 s for 6 M points)."""
    lib = _lib.load()
    device = torch.device(device)
    mask = _lib.empty(n, dtype=torch.bool, device=device)
    nbytes = C.c_size_t(0)
    _lib.check(lib.cnrma_sample_workspace_bytes(C.byref(nbytes)), "cnrma_sample_workspace_bytes")
    ws = _lib.empty(nbytes.value, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.cnrma_sample_mask(n, int(max_points), int(seed) & 0xFFFFFFFFFFFFFFFF, C.c_void_p(ws.data_ptr()),
                                         nbytes.value, C.c_void_p(mask.data_ptr()), _stream(device)), "cnrma_sample_mask")
    mask._cnrma_count = min(int(n), int(max_points))     # exact-k draw: see _mask_count
    return mask


def _mask_prefix(mask_dev):
    lib = _lib.load()
    device = mask_dev.device
    n = mask_dev.numel()
    nbytes = C.c_size_t(0)
    _lib.check(lib.cnrma_handoff_workspace_bytes(n, C.byref(nbytes)), "cnrma_handoff_workspace_bytes")
    ws = _lib.empty(nbytes.value, dtype=torch.uint8, device=device)
    prefix = _lib.empty(n, dtype=torch.int32, device=device)
    kept = _lib.empty(1, dtype=torch.int64, device=device)
    _lib.check(lib.cnrma_mask_prefix(C.c_void_p(mask_dev.data_ptr()), n, C.c_void_p(ws.data_ptr()), nbytes.value,
                                     C.c_void_p(prefix.data_ptr()), C.c_void_p(kept.data_ptr()), _stream(device)),
               "cnrma_mask_prefix")
    return prefix, kept


def _mask_count(mask):
    """Number of kept rows.  Masks drawn by sample_points_device carry it (exactly min(n, max_points) by
    construction), which saves a device reduction and a host synchronisation."""
    known = getattr(mask, "_cnrma_count", None)
    return int(known) if known is not None else int(mask.sum())


def _to_mask_dev(mask, n, device):
    if isinstance(mask, torch.Tensor):
        m = mask.to(device=device, dtype=torch.bool)
    else:
        m = torch.from_numpy(mask.astype(bool)).to(device)
    if m.numel() != n:
        raise ValueError("mask length must equal the number of points")
    return m.contiguous()


class _SelectRowsBackward(torch.autograd.Function):
    """Attaches the gradient of the ordered row selection to its result: kept row j came from row idx[j] of `points`,
    `coord + offset` has gradient one (rm.py:362-402 are differentiable in the reference: the detection loss reaches the
    2D features through the selected points)."""

    @staticmethod
    def forward(ctx, out, mask_dev, points):
        ctx.save_for_backward(mask_dev)
        ctx.shape = tuple(points.shape)
        return out.view_as(out)

    @staticmethod
    def backward(ctx, grad_out):
        (mask_dev,) = ctx.saved_tensors
        grad = torch.zeros(ctx.shape, dtype=grad_out.dtype, device=grad_out.device)
        grad.masked_scatter_(mask_dev.unsqueeze(1).expand(ctx.shape), grad_out.contiguous())   # row order preserved
        return None, None, grad


def switch_pointcloud(points, offsets, max_points=None, masks=None, rng=None):
    """The tensor part of RayMarching.switch_pointcloud (rm.py:339-407; augmentation and boxes stay with the caller):
    per batch element `coord + offset`, then the rows kept by sample_points' mask, in order.

    points: list of [N,3+C] CUDA tensors; offsets: list of 3-vectors; masks: optional list of bool masks (else drawn
    like the reference when max_points is set).  Returns (coords list of [Nsel,3], features list of [Nsel,C])."""
    lib = _lib.load()
    coords, feats = [], []
    for b, pts in enumerate(points):
        if not pts.is_cuda or pts.dtype != torch.float32:
            raise CnrmaError("points must be float32 CUDA tensors")
        device = pts.device
        n, cols = pts.shape
        off = _origin3(offsets[b])
        mask = masks[b] if masks is not None else (sample_points(n, max_points, rng) if max_points is not None else None)
        if mask is None:
            mask_dev = torch.ones(n, dtype=torch.bool, device=device)
            n_sel = n
        else:
            n_sel = _mask_count(mask)
            mask_dev = _to_mask_dev(mask, n, device)
        out = _lib.empty((n_sel, cols), dtype=torch.float32, device=device)
        if n_sel > 0:
            with torch.cuda.device(device):
                prefix, _kept = _mask_prefix(mask_dev)
                off3 = (C.c_float * 3)(*off)
                src = pts if pts.stride(1) == 1 else pts.contiguous()
                src = src.detach()
                _lib.check(lib.cnrma_select_rows(C.c_void_p(src.data_ptr()), src.stride(0), cols, n,
                                                 C.c_void_p(mask_dev.data_ptr()), C.c_void_p(prefix.data_ptr()), off3,
                                                 C.c_void_p(out.data_ptr()), cols, n_sel, _stream(device)),
                           "cnrma_select_rows")
        if torch.is_grad_enabled() and pts.requires_grad:
            out = _SelectRowsBackward.apply(out, mask_dev, pts)
        coords.append(out[:, 0:3])
        feats.append(out[:, 3:])
    return coords, feats


def _rows_view(coords, feats):
    """[N,3] coords and [N,C] features as one [N,3+C] row tensor: the buffer they are both views of when they come from
    switch_pointcloud / rma_points_selected, else a concatenation."""
    base = coords._base
    if (base is not None and feats._base is base and base.dim() == 2 and base.is_contiguous()
            and base.shape[1] == 3 + feats.shape[1] and coords.data_ptr() == base.data_ptr()
            and feats.data_ptr() == base.data_ptr() + 3 * base.element_size() and coords.shape[0] <= base.shape[0]):
        return base[: coords.shape[0]]
    return torch.cat((coords, feats), dim=1).contiguous()


def quantize_points(coords, feats, voxel_size, batch_index=None):
    """The data side of `ME.utils.batch_sparse_collate` + `ME.SparseTensor` for ONE batch element (rm.py:330-332;
    MinkowskiEngine 0.5.4 is not part of the reference tree): cells = trunc(coords / voxel_size) as int32, one row per
    occupied cell -- the first in row order, survivors in row order (MinkowskiEngine keeps an arbitrary duplicate; this
    is the deterministic choice).  coords [N,3] float32 CUDA (the reference passes the un-divided coordinates and
    divides in the call, rm.py:331), feats [N,C].
    Returns (cells int32 [K,3] -- or [K,4] with `batch_index` in column 0, as sparse_collate lays them out --,
    features [K,C], coords [K,3])."""
    lib = _lib.load()
    if not coords.is_cuda or coords.dtype != torch.float32 or feats.dtype != torch.float32:
        raise CnrmaError("coords / feats must be float32 CUDA tensors")
    if torch.is_grad_enabled() and (coords.requires_grad or feats.requires_grad):
        raise CnrmaError("quantize_points has no backward: quantise detached tensors (MinkowskiEngine re-attaches the "
                         "features it keeps through its own index)")
    device = coords.device
    rows = _rows_view(coords.detach(), feats.detach())
    n, cols = rows.shape
    lead = 0 if batch_index is None else 1
    if n == 0:
        return (torch.zeros((0, 3 + lead), dtype=torch.int32, device=device), rows[:, 3:], rows[:, :3])
    with torch.cuda.device(device):
        nbytes = C.c_size_t(0)
        _lib.check(lib.cnrma_quantize_workspace_bytes(n, C.byref(nbytes)), "cnrma_quantize_workspace_bytes")
        ws = _lib.empty(nbytes.value, dtype=torch.uint8, device=device)
        keep = _lib.empty(n, dtype=torch.bool, device=device)
        _lib.check(lib.cnrma_quantize_mark(C.c_void_p(rows.data_ptr()), rows.stride(0), n, float(voxel_size),
                                           C.c_void_p(ws.data_ptr()), nbytes.value, C.c_void_p(keep.data_ptr()),
                                           _stream(device)), "cnrma_quantize_mark")
        prefix, kept = _mask_prefix(keep)
        k = int(kept.item())                                    # the one host sync: the number of occupied cells
        out = _lib.empty((k, cols), dtype=torch.float32, device=device)
        cells = _lib.empty((k, 3), dtype=torch.int32, device=device)
        if k > 0:
            _lib.check(lib.cnrma_quantize_compact(C.c_void_p(rows.data_ptr()), rows.stride(0), cols, n, float(voxel_size),
                                                  C.c_void_p(keep.data_ptr()), C.c_void_p(prefix.data_ptr()),
                                                  C.c_void_p(out.data_ptr()), cols, C.c_void_p(cells.data_ptr()), k,
                                                  _stream(device)), "cnrma_quantize_compact")
    if batch_index is not None:
        cells = torch.cat((torch.full((k, 1), int(batch_index), dtype=torch.int32, device=device), cells), dim=1)
    return cells, out[:, 3:], out[:, :3]


def sparse_collate_quantized(coords_list, feats_list, voxel_size):
    """Drop-in for the data preparation of rm.py:330-332 over the batch: returns (coordinates int32 [K,4] with the batch
    index in column 0, features [K,C]) ready for `ME.SparseTensor(coordinates=..., features=...)`, already unique."""
    cells, feats = [], []
    for b, (c, f) in enumerate(zip(coords_list, feats_list)):
        q, ff, _c = quantize_points(c, f, voxel_size, batch_index=b)
        cells.append(q)
        feats.append(ff)
    return torch.cat(cells, dim=0), torch.cat(feats, dim=0)


def _selected_without_sync(lib, fs, m, grid, desc, cap, max_points, seed, off, cols, device):
    """Sampler (row count from the march's result block on the device), prefix sum and selected fill for at most `cap`
    rows, queued without reading M: returns the [max_points, cols] buffer whose first min(M, max_points) rows are valid
    provided M <= cap."""
    mask = _lib.empty(cap, dtype=torch.bool, device=device)
    nbytes = C.c_size_t(0)
    _lib.check(lib.cnrma_sample_workspace_bytes(C.byref(nbytes)), "cnrma_sample_workspace_bytes")
    ws = _lib.empty(nbytes.value, dtype=torch.uint8, device=device)
    _lib.check(lib.cnrma_sample_mask_for_result(C.c_void_p(m.result.data_ptr()), cap, max_points,
                                                seed & 0xFFFFFFFFFFFFFFFF, C.c_void_p(ws.data_ptr()), nbytes.value,
                                                C.c_void_p(mask.data_ptr()), _stream(device)), "cnrma_sample_mask_for_result")
    prefix, _kept = _mask_prefix(mask)
    out = _lib.empty((max_points, cols), dtype=torch.float32, device=device)
    off3 = (C.c_float * 3)(*off)
    _lib.check(lib.cnrma_rma_fill_selected(
        C.byref(grid), C.c_void_p(m.pinv.data_ptr()), C.byref(desc), m.grids, m.t_one, m.mode, m.threshold,
        m.depth_points, C.c_void_p(m.workspace.data_ptr()), C.c_void_p(m.result.data_ptr()), 1, None,
        C.c_void_p(mask.data_ptr()), C.c_void_p(prefix.data_ptr()), cap, off3, C.c_void_p(out.data_ptr()), cols, max_points,
        _stream(device)), "cnrma_rma_fill_selected")
    return out


def rma_points_selected(projections, features, tsdf, voxel_dim, voxel_size, origin, stride, offsets, max_points=None,
                        masks=None, rng=None, grids=300, mode="neus", threshold=None, depth_points=None,
                        device_seed=None):
    """aggregate_2d_features_ray_marching (rm.py:260-307) fused with switch_pointcloud (rm.py:339-407): the march
    runs as usual, M is read back, the keep mask is drawn (or taken from `masks`: per batch element a mask or a
    callable `n -> mask`), and the fill kernel produces ONLY
    the kept rows, offset already added -- the un-sampled point cloud (M x (3+C) floats) is never written.

    `device_seed` (with `max_points`): the mask is drawn on the device like sample_points_device, from the row count
    the march leaves IN DEVICE MEMORY -- sampler, prefix sum and fill are queued right behind the march and M is read
    back afterwards, so the host never stalls between the kernels (buffers are sized from recent calls with the same
    shapes; the first call, or a scene that keeps more rows than any recent one, takes the path above).

    Inference only: raises when a feature map requires grad (train through rma_points + switch_pointcloud, whose
    results carry the backward kernels).
    Returns (coords list of [Nsel,3], features list of [Nsel,C]) like switch_pointcloud."""
    _check_mode(mode, threshold, depth_points)
    lib = _lib.load()
    if _needs_grad([features] if isinstance(features, torch.Tensor) else _as_view_list(features)):
        raise CnrmaError("rma_points_selected has no backward: use rma_points + switch_pointcloud when training")
    fs = _FeatureStack(features, need_vector_layout=False)
    device = fs.device
    if not isinstance(projections, torch.Tensor):
        projections = torch.stack(list(projections), dim=0)
    P_scaled = scale_projections(projections, stride)
    grid = _lib.make_grid(voxel_dim, voxel_size, _origin3(origin))
    coords, feats = [], []
    with torch.cuda.device(device):
        for b in range(fs.B):
            m = _march(fs, b, P_scaled[:, b], tsdf[b, 0], grid, voxel_dim, voxel_size, grids, mode, threshold,
                       depth_points)
            desc = fs.descriptor(b)
            if _phase_hook is not None:
                _phase_hook("march")
            cols = fs.C + 3
            key = ("selected", device.index, b, fs.V, fs.C, fs.H, fs.W, tuple(int(v) for v in voxel_dim), int(grids), mode,
                   threshold, depth_points)
            if device_seed is not None and max_points is not None and masks is None:
                cap = _hint_rows(key)
                if cap:
                    out = _selected_without_sync(lib, fs, m, grid, desc, cap, int(max_points), int(device_seed) + b,
                                                 _origin3(offsets[b]), cols, device)
                    n = int(_read_result(m).rows)             # everything is queued: the host waits here only
                    _note_rows(key, n)
                    if n <= cap:
                        out = out[: min(n, int(max_points))]
                        coords.append(out[:, 0:3])
                        feats.append(out[:, 3:])
                        continue
                else:
                    n = int(_read_result(m).rows)
                    _note_rows(key, n)
                mask = sample_points_device(n, max_points, int(device_seed) + b, device)
            else:
                n = int(_read_result(m).rows)
                mask = masks[b] if masks is not None else (sample_points(n, max_points, rng) if max_points is not None
                                                           else None)
            if callable(mask):                      # drawn once M is known, e.g. lambda n: sample_points_device(n, ...)
                mask = mask(n)
            if mask is None:
                mask_dev, n_sel = torch.ones(n, dtype=torch.bool, device=device), n
            else:
                n_sel = _mask_count(mask)
                mask_dev = _to_mask_dev(mask, n, device)
            out = _lib.empty((n_sel, cols), dtype=torch.float32, device=device)
            if n_sel > 0:
                prefix, _kept = _mask_prefix(mask_dev)
                off3 = (C.c_float * 3)(*_origin3(offsets[b]))
                _lib.check(lib.cnrma_rma_fill_selected(
                    C.byref(grid), C.c_void_p(m.pinv.data_ptr()), C.byref(desc), m.grids, m.t_one, m.mode, m.threshold,
                    m.depth_points, C.c_void_p(m.workspace.data_ptr()), C.c_void_p(m.result.data_ptr()), 1, None,
                    C.c_void_p(mask_dev.data_ptr()), C.c_void_p(prefix.data_ptr()), n, off3, C.c_void_p(out.data_ptr()),
                    cols, n_sel, _stream(device)), "cnrma_rma_fill_selected")
            coords.append(out[:, 0:3])
            feats.append(out[:, 3:])
    return coords, feats
