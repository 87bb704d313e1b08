"""TEST INFRASTRUCTURE ONLY -- import shim for the *real* CN-RMA reference.

Usable where /root/reference is mounted (the build container) or where `oracle/build_ref.py` has staged a copy of
the reference's Python package under oracle/_ref (git-ignored; it travels to the GPU box with the snapshot, which is
how bench.py's CPU legs time the reference itself there).  It lets
`oracle/make_golden.py`, `tests/test_oracle_vs_reference.py` and `oracle/ref_runner.py` import the
reference's `projects/mvsdetection/models/ray_marching.py` unmodified by
registering stub modules for the third-party packages that file imports at
module scope (mmcv / mmdet / mmdet3d / MinkowskiEngine / open3d / skimage /
trimesh) -- none of which the hot path (ray_marching.py:21-111, :200-307,
:687-956) actually uses.  Nothing here is on the product path.
"""
import importlib
import os
import sys
import types
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("CNRMA_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "projects", "mvsdetection")):
            return cand
    return os.environ.get("CNRMA_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()

_STUBS = [
    "open3d", "MinkowskiEngine", "MinkowskiEngine.modules", "MinkowskiEngine.modules.resnet_block",
    "mmdet", "mmdet.models", "mmdet.models.builder", "mmdet.datasets", "mmdet.datasets.builder",
    "mmdet.core", "mmdet.core.bbox", "mmdet.core.bbox.builder", "mmdet3d", "mmdet3d.core",
    "mmdet3d.core.bbox", "mmdet3d.core.bbox.structures", "mmdet3d.ops", "mmdet3d.ops.pcdet_nms",
    "mmdet3d.datasets", "mmcv", "mmcv.runner", "mmcv.parallel", "mmcv.cnn", "skimage",
    "skimage.measure", "trimesh",
]


class _Registry:
    """Stands in for an mmcv Registry: register_module works bare or called."""

    def register_module(self, cls=None, **_kw):
        if cls is not None and isinstance(cls, type):
            return cls
        return lambda c: c


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        val = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, val)
        return val


def _identity_decorator_factory(*_a, **_kw):
    return lambda fn: fn


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "projects", "mvsdetection"))


_cached = None


def load_reference():
    """Returns the reference's ray_marching module (imported from REFERENCE_ROOT)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    for mod, names in {
        "mmdet.models": ["DETECTORS", "BACKBONES", "HEADS", "NECKS"],
        "mmdet.models.builder": ["DETECTORS", "BACKBONES", "HEADS", "NECKS"],
        "mmdet.datasets": ["DATASETS", "PIPELINES"],
        "mmdet.datasets.builder": ["DATASETS", "PIPELINES"],
        "mmdet3d.datasets": ["DATASETS", "PIPELINES"],
        "mmdet.core.bbox.builder": ["BBOX_ASSIGNERS"],
    }.items():
        for n in names:
            setattr(sys.modules[mod], n, _Registry())
    # names used as base classes must be real classes
    sys.modules["mmdet3d.datasets"].Custom3DDataset = type("Custom3DDataset", (), {})
    sys.modules["mmdet.core"].BaseAssigner = type("BaseAssigner", (), {})
    sys.modules["mmdet.core.bbox"].BaseAssigner = sys.modules["mmdet.core"].BaseAssigner
    rb = sys.modules["MinkowskiEngine.modules.resnet_block"]
    rb.BasicBlock = type("BasicBlock", (), {})
    rb.Bottleneck = type("Bottleneck", (), {})
    sys.modules["mmcv.runner"].auto_fp16 = _identity_decorator_factory
    sys.modules["mmcv.runner"].force_fp32 = _identity_decorator_factory
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _cached = importlib.import_module("projects.mvsdetection.models.ray_marching")
    return _cached


def make_self(voxel_dim, voxel_size, origin, stride=4, ray_marching_type="neus",
              neus_threshold=0.05, depth_points=None):
    """A bare `self` carrying the attributes the reference's hot methods read
    (ray_marching.py:166-194), with the unbound methods attached."""
    import torch
    rm = load_reference()
    s = types.SimpleNamespace()
    s.voxel_dim = tuple(voxel_dim)
    s.voxel_size = voxel_size
    s.origin = origin if isinstance(origin, torch.Tensor) else torch.tensor(origin).view(1, 3)
    s.backbone2d_stride = stride
    s.ray_marching_type = ray_marching_type
    s.neus_threshold = neus_threshold
    s.depth_points = depth_points
    for name in ("initialize_volume", "aggregate_2d_features", "clear_3d_features",
                 "aggregate_2d_features_ray_marching", "ray_projection_neus", "ray_projection_depth"):
        setattr(s, name, types.MethodType(getattr(rm.RayMarching, name), s))
    s.initialize_volume()
    return s
