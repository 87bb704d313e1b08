"""Host mirror of the reference's `TSDFFusion` (data_prepare/scannet/tsdf.py:353-475, same class in
data_prepare/arkit/tsdf.py): same constructor arguments, `reset`, `integrate(projection, depth, color, label)` and
`get_tsdf`, backed by cnrma_tsdf_integrate.  `integrate_frames` fuses a whole batch of frames in one pass over the
volume (frame order preserved, so the result is bit-identical to frame-by-frame calls).

`get_tsdf` returns the normalised volumes as tensors (tsdf [nx,ny,nz], {'color': [3,nx,ny,nz], label_name: [nx,ny,nz]});
the reference wraps the same tensors in its `TSDF` container (mesh extraction via skimage / trimesh, out of scope).
"""
import ctypes as C

import torch

from . import _lib
from .functional import _origin3, _stream


class TSDFFusion:
    def __init__(self, voxel_dim=(128, 128, 128), voxel_size=.02, origin=(0, 0, 0), trunc_ratio=3,
                 device=torch.device('cuda'), color=True, label=False):
        nx, ny, nz = (int(v) for v in voxel_dim)
        self.voxel_dim = (nx, ny, nz)
        self.voxel_size = voxel_size
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.CnrmaError("TSDFFusion needs a CUDA device: there is no CPU path")
        self.device = device
        self.origin = torch.tensor(_origin3(origin), dtype=torch.float, device=device).view(1, 3)
        self.trunc_margin = voxel_size * trunc_ratio                      # tsdf.py:373
        n = nx * ny * nz
        self.tsdf_vol = torch.ones(n, device=device)
        self.weight_vol = torch.zeros(n, device=device)
        self.color_vol = torch.zeros((3, n), device=device) if color else None
        self.label_vol = -torch.ones(n, device=device, dtype=torch.long) if label else None
        self._grid = _lib.make_grid(self.voxel_dim, voxel_size, _origin3(origin))

    def reset(self):
        """Initialize the volumes to default values (tsdf.py:393-400)."""
        self.tsdf_vol.fill_(1)
        self.weight_vol.fill_(0)
        if self.color_vol is not None:
            self.color_vol.fill_(0)
        if self.label_vol is not None:
            self.label_vol.fill_(-1)

    def integrate(self, projection, depth, color=None, label=None):
        """Accumulate one depth map (and colour / label image) into the TSDF (tsdf.py:402-451)."""
        self.integrate_frames(projection.unsqueeze(0), depth.unsqueeze(0),
                              None if color is None else color.unsqueeze(0),
                              None if label is None else label.unsqueeze(0))

    def integrate_frames(self, projections, depths, colors=None, labels=None):
        """projections [F,3,4], depths [F,H,W], colors [F,3,H,W] or None, labels [F,H,W] int64 or None."""
        lib = _lib.load()
        dev = self.device
        P = projections.detach().to(device=dev, dtype=torch.float32).contiguous()
        D = depths.detach().to(device=dev, dtype=torch.float32).contiguous()
        F_, H, W = D.shape
        Cimg = None if (colors is None or self.color_vol is None) else colors.detach().to(device=dev, dtype=torch.float32).contiguous()
        L = None if (labels is None or self.label_vol is None) else labels.detach().to(device=dev, dtype=torch.long).contiguous()

        def table(t):
            if t is None:
                return None
            return (C.c_void_p * F_)(*[t[i].data_ptr() for i in range(F_)])
        dt, ct, lt = table(D), table(Cimg), table(L)
        with torch.cuda.device(dev):
            _lib.check(lib.cnrma_tsdf_integrate(
                C.byref(self._grid), C.c_void_p(P.data_ptr()), 12, F_, dt, ct, lt, H, W, float(self.trunc_margin),
                C.c_void_p(self.tsdf_vol.data_ptr()), C.c_void_p(self.weight_vol.data_ptr()),
                C.c_void_p(self.color_vol.data_ptr()) if (self.color_vol is not None and Cimg is not None) else None,
                C.c_void_p(self.label_vol.data_ptr()) if (self.label_vol is not None and L is not None) else None,
                _stream(dev)), "cnrma_tsdf_integrate")

    def get_tsdf(self, label_name='instance'):
        """The normalised volumes (tsdf.py:453-475): tsdf / weight and colour / weight where weight > 0."""
        nx, ny, nz = self.voxel_dim
        seen = self.weight_vol > 0
        tsdf_vol = self.tsdf_vol.clone()
        tsdf_vol[seen] /= self.weight_vol[seen]
        attribute_vols = {}
        if self.color_vol is not None:
            color_vol = self.color_vol.clone()
            color_vol[:, seen] /= self.weight_vol[seen]
            attribute_vols['color'] = color_vol.view(3, nx, ny, nz)
        if self.label_vol is not None:
            attribute_vols[label_name] = self.label_vol.view(nx, ny, nz).clone()
        return tsdf_vol.view(nx, ny, nz), attribute_vols
