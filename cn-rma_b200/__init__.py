"""cn-rma_b200: B200-native (sm_100a) ray-marching aggregation for CN-RMA -- the 2D -> 3D feature lift of
projects/mvsdetection/models/ray_marching.py behind the reference's own function names.

Import name: `cnrma_b200` (see the alias package at the repository root; '-' is not valid in a module name).
"""
from ._lib import CnrmaError, build, load, reload_tuning, LIB_PATH, EXPORTS  # noqa: F401
from . import _lib  # noqa: F401
from .functional import (aggregate_views, aggregate_views_bilinear, backproject, dense_rma, finalize_views, get_ray_parameter,  # noqa: F401
                         invert_projections, project_views, rma_points_selected, sample_points, sample_points_device,
                         switch_pointcloud, ray_projection, rma_dense_weights, rma_points, scale_projections,
                         quantize_points, sparse_collate_quantized, set_phase_hook, fill_stats)
from .module import RayMarchingAggregator, make_atlas_class, make_detector_class  # noqa: F401
from .fusion import TSDFFusion  # noqa: F401
from .tsdf_head import AtlasTSDFHead, tsdf_head_scale  # noqa: F401
from . import distributed, synthetic  # noqa: F401

__all__ = ["CnrmaError", "build", "load", "aggregate_views", "backproject", "dense_rma", "get_ray_parameter",
           "project_views", "ray_projection", "rma_dense_weights", "rma_points", "RayMarchingAggregator",
           "make_detector_class", "synthetic"]
