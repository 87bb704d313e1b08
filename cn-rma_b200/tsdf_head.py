"""Host mirror of the reference's `AtlasTSDFHead` (projects/mvsdetection/models/atlas_head.py:15-82, "ah.py"): same
constructor arguments, same parameter names (`decoders.{i}.weight`, shape [1,C,1,1,1] -- a reference checkpoint loads
with `load_state_dict` as is), same `forward(xs, targets=None) -> (output, losses)` with the same dictionary keys.

The per-scale work of ah.py:38-52 (1x1x1 convolution, tanh, label smoothing, nearest x2 upsampling of the previous
scale, sparsification) is one call of cnrma_tsdf_head_scale per scale; its autograd (gradients to `xs` and to the
decoder weights) is cnrma_tsdf_head_scale_backward.  The loss bookkeeping of ah.py:57-82 is a handful of reductions on
the outputs and stays in torch.

`output['scene_tsdf_004']` is what `aggregate_2d_features_ray_marching` takes as `tsdf` (rm.py:440, :490).
"""
import ctypes as C

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib
from .functional import _stream

_DTYPES = {torch.float32: 0, torch.bfloat16: 1}


def _strides(x):
    """(stride_c, stride_v) of one batch element [C,X,Y,Z] whose voxels are flat-addressable, else None."""
    _c, nx, ny, nz = x.shape
    sc, sx, sy, sz = x.stride()
    if sy == nz * sz and sx == ny * nz * sz:
        return sc, sz
    return None


def _flat(x):
    """x [B,C,X,Y,Z] in a layout the kernel addresses directly (NCDHW or channels_last_3d), else a contiguous copy."""
    if x.dim() != 5:
        raise ValueError("expected [B,C,X,Y,Z]")
    if _strides(x[0]) is None:
        x = x.contiguous()
    return x


def _scale_forward(x, weight, prev, label_smoothing, sparse_threshold, want_mask):
    lib = _lib.load()
    if not x.is_cuda:
        raise _lib.CnrmaError("the TSDF head needs CUDA tensors: there is no CPU path")
    if x.dtype not in _DTYPES:
        raise _lib.CnrmaError(f"unsupported feature dtype {x.dtype}")
    x = _flat(x)
    B, Cc, nx, ny, nz = x.shape
    w = weight.detach().reshape(-1).to(device=x.device, dtype=torch.float32).contiguous()
    if w.numel() != Cc:
        raise ValueError(f"decoder has {w.numel()} input channels, the volume {Cc}")
    if prev is not None:
        prev = prev.detach().to(device=x.device, dtype=torch.float32).contiguous()
        if tuple(prev.shape) != (B, 1, nx // 2, ny // 2, nz // 2) or (nx | ny | nz) & 1:
            raise ValueError(f"previous scale {tuple(prev.shape)} is not half of {tuple(x.shape)}")
    tsdf = _lib.empty((B, 1, nx, ny, nz), dtype=torch.float32, device=x.device)
    mask = _lib.empty((B, 1, nx, ny, nz), dtype=torch.bool, device=x.device) if want_mask else None
    with torch.cuda.device(x.device):
        for b in range(B):
            sc, sv = _strides(x[b])
            _lib.check(lib.cnrma_tsdf_head_scale(
                C.c_void_p(x[b].data_ptr()), _DTYPES[x.dtype], Cc, nx, ny, nz, sc, sv, C.c_void_p(w.data_ptr()),
                C.c_void_p(prev[b].data_ptr()) if prev is not None else None, float(label_smoothing),
                float(sparse_threshold) if prev is not None else 0.0, C.c_void_p(tsdf[b].data_ptr()),
                C.c_void_p(mask[b].data_ptr()) if mask is not None else None, _stream(x.device)), "cnrma_tsdf_head_scale")
    return x, w, prev, tsdf, mask


class _HeadScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, prev, label_smoothing, sparse_threshold):
        xf, w, p, tsdf, mask = _scale_forward(x, weight, prev, label_smoothing, sparse_threshold, True)
        ctx.save_for_backward(xf, w, p if p is not None else torch.empty(0, device=x.device), tsdf)
        ctx.has_prev = p is not None
        ctx.args = (float(label_smoothing), float(sparse_threshold) if p is not None else 0.0, weight.shape)
        ctx.mark_non_differentiable(mask)
        return tsdf, mask

    @staticmethod
    def backward(ctx, grad_tsdf, _grad_mask):
        lib = _lib.load()
        x, w, prev, tsdf = ctx.saved_tensors
        ls, thr, wshape = ctx.args
        if x.dtype != torch.float32:
            raise _lib.CnrmaError("the TSDF head backward takes float32 volumes")
        B, Cc, nx, ny, nz = x.shape
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g = grad_tsdf.detach().to(torch.float32).contiguous()
        grad_x = _lib.empty_like(x) if need_x else None           # same strides as x
        grad_w = torch.zeros(Cc, dtype=torch.float32, device=x.device) if need_w else None
        gw_b = _lib.empty(Cc, dtype=torch.float32, device=x.device) if need_w else None
        ws = None
        if need_w:
            ws = _lib.empty(int(lib.cnrma_tsdf_head_workspace_bytes(Cc)), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            for b in range(B):
                sc, sv = _strides(x[b])
                _lib.check(lib.cnrma_tsdf_head_scale_backward(
                    C.c_void_p(x[b].data_ptr()), Cc, nx, ny, nz, sc, sv, C.c_void_p(w.data_ptr()),
                    C.c_void_p(prev[b].data_ptr()) if ctx.has_prev else None, C.c_void_p(tsdf[b].data_ptr()),
                    C.c_void_p(g[b].data_ptr()), ls, thr, C.c_void_p(grad_x[b].data_ptr()) if need_x else None,
                    C.c_void_p(gw_b.data_ptr()) if need_w else None, C.c_void_p(ws.data_ptr()) if need_w else None,
                    ws.numel() if need_w else 0, _stream(x.device)), "cnrma_tsdf_head_scale_backward")
                if need_w:
                    grad_w += gw_b
        return grad_x, (grad_w.view(wshape) if need_w else None), None, None, None


def tsdf_head_scale(x, weight, prev=None, label_smoothing=1.05, sparse_threshold=0.99):
    """One scale of ah.py:38-52: x [B,C,X,Y,Z], weight [1,C,1,1,1] (or [C]), prev [B,1,X/2,Y/2,Z/2] or None ->
    (tsdf [B,1,X,Y,Z] f32, surface mask [B,1,X,Y,Z] bool).  Differentiable w.r.t. x and weight."""
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        return _HeadScale.apply(x, weight, prev, label_smoothing, sparse_threshold)
    _x, _w, _p, tsdf, mask = _scale_forward(x, weight, prev, label_smoothing, sparse_threshold, True)
    return tsdf, mask


def log_transform(x, shift=1):
    """ah.py:84-87."""
    return x.sign() * (1 + x.abs() / shift).log()


class AtlasTSDFHead(nn.Module):
    """Drop-in for the reference head (ah.py:15-82)."""

    def __init__(self, input_channels, n_scales, voxel_size, label_smoothing, sparse_threshold):
        super().__init__()
        self.fp16_enabled = False
        self.input_channels = input_channels
        self.n_scales = n_scales
        self.voxel_size = voxel_size
        self.label_smoothing = label_smoothing
        self.sparse_threshold = sparse_threshold
        self.voxel_sizes = [self.voxel_size * (2 ** i) for i in range(n_scales)][::-1]
        self.keys = [str(int(voxel_size * 100)).zfill(3) for voxel_size in self.voxel_sizes]
        # parameter holders only (same names / shapes / init as ah.py:29-31); the convolution runs in the fused kernel
        self.decoders = nn.ModuleList([nn.Conv3d(c, 1, 1, bias=False) for c in self.input_channels][::-1])

    def forward(self, xs, targets=None):
        output, losses, mask_surface_pred = {}, {}, []
        prev = None
        for i, (decoder, x) in enumerate(zip(self.decoders, xs)):
            tsdf, mask = tsdf_head_scale(x.float() if x.dtype == torch.float16 else x, decoder.weight, prev,
                                         self.label_smoothing, self.sparse_threshold[i - 1] if i > 0 else 0.0)
            if i > 0:
                mask_surface_pred.append(mask)
            output['scene_tsdf_' + self.keys[i]] = tsdf
            prev = tsdf
        if targets is not None:
            losses = self.losses(output, mask_surface_pred, targets)
        return output, losses

    def losses(self, output, mask_surface_pred, targets):
        """ah.py:57-82: L1 between log-transformed prediction and target over observed (or wholly outside) voxels,
        restricted from the second scale on to the voxels the previous scale predicted as surface."""
        losses = {}
        for i, key in enumerate(self.keys):
            pred = output['scene_tsdf_' + key]
            trgt = targets['tsdf_gt_' + key]
            keep = (trgt < 1) | (trgt == 1).all(-1, keepdim=True)
            if i > 0:
                keep = mask_surface_pred[i - 1] & keep
            loss = F.l1_loss(log_transform(pred, 1.0), log_transform(trgt, 1.0), reduction='none')
            if i == 0 or keep.sum() > 0:
                losses['tsdf_loss_' + key] = loss[keep].mean()
            else:
                losses['tsdf_loss_' + key] = 0 * loss.sum()
        return losses
