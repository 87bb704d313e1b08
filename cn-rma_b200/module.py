"""Stateful mirror of the aggregation half of the reference's `RayMarching` detector
(projects/mvsdetection/models/ray_marching.py, "rm.py"): same attribute and method names, same argument
meaning, so the detector's forward_train / forward_test (rm.py:409-521) run unchanged on top of it.

    self.initialize_volume()                                              rm.py:200-209
    self.aggregate_2d_features(projection, feature)                       rm.py:220-244
    self.clear_3d_features()                                              rm.py:247-257
    self.aggregate_2d_features_ray_marching(projections, features, tsdf)  rm.py:260-307
    self.ray_projection_neus(projection, features, tsdf, grids, weight_threshold)   rm.py:687
    self.ray_projection_depth(projection, features, tsdf, grids, select_grids)      rm.py:809
    state: volume, valid, points_detection, voxel_dim, voxel_size, origin, backbone2d_stride,
           ray_marching_type, neus_threshold, depth_points

The one structural difference: `aggregate_2d_features` only records its (projection, feature) pair; the
fused kernel runs when `clear_3d_features()` is called or when `volume` / `valid` are read, visiting the
views in call order so the fp32 sums are bit-identical to the reference's per-view `self.volume + volume`.
"""
import torch

from . import functional as F


class RayMarchingAggregator:
    def __init__(self, voxel_size, voxel_dim, origin=(0, 0, 0), backbone2d_stride=4, ray_marching_type="neus",
                 neus_threshold=None, depth_points=None, max_points=None, feature_transform=None):
        self.voxel_size = voxel_size
        self.voxel_dim = tuple(voxel_dim)
        self.origin = torch.tensor(origin).view(1, 3)         # rm.py:184
        self.backbone2d_stride = backbone2d_stride
        self.ray_marching_type = ray_marching_type
        self.neus_threshold = neus_threshold
        self.depth_points = depth_points
        self.max_points = max_points                  # rm.py:182
        self.feature_transform = feature_transform    # rm.py:158-161 (augmentation: the caller's callable)
        if ray_marching_type == "neus":                        # rm.py:191-194
            assert neus_threshold is not None
        elif ray_marching_type == "depth":
            assert depth_points in [0, 1, 2, 3, 4]
        self.initialize_volume()

    # ---- Stage A ------------------------------------------------------------------------------------
    def initialize_volume(self):
        """Reset the accumulators (rm.py:200-209)."""
        self._pending = []
        self._sum = None          # (volume_sum, count_i32, valid_bool) once views have been folded in
        self._cleared = None      # (volume_mean, valid_bool) after clear_3d_features
        self.points_detection = []

    def aggregate_2d_features(self, projection, feature):
        """projection [B,3,4] un-scaled, feature [B,C,H',W'] (rm.py:220-244).  Deferred: see module docstring."""
        if self._cleared is not None:
            raise RuntimeError("aggregate_2d_features after clear_3d_features: call initialize_volume first")
        self._pending.append((projection, feature))

    def _flush(self, mean):
        if self._pending:
            projections = torch.stack([p for p, _ in self._pending], dim=0)
            features = [f for _, f in self._pending]
            self._sum = F.aggregate_views(projections, features, self.voxel_dim, self.voxel_size, self.origin,
                                          self.backbone2d_stride, mean=mean, out=self._sum)
            self._pending = []
        elif mean and self._sum is not None:
            raise RuntimeError("internal: mean requested after the sums were materialised")

    def clear_3d_features(self):
        """Mean over the views that see each voxel, 0 elsewhere; valid becomes bool (rm.py:247-257)."""
        if self._sum is not None and self._pending:
            self._flush(mean=True)      # accumulate the rest into the existing sums, then divide
        elif self._sum is not None:
            vol, cnt, valid = self._sum   # sums were already read back un-averaged: finish them in place
            F.finalize_views(vol, cnt, valid)
        else:
            self._flush(mean=True)
        if self._sum is None:
            raise RuntimeError("clear_3d_features without any aggregated view")
        vol, _cnt, valid = self._sum
        self._cleared = (vol, valid)

    @property
    def volume(self):
        if self._cleared is not None:
            return self._cleared[0]
        self._flush(mean=False)
        return 0 if self._sum is None else self._sum[0]

    @property
    def valid(self):
        """View count (int64, like the reference's `0 + bool` sums) before clear_3d_features, bool after."""
        if self._cleared is not None:
            return self._cleared[1]
        self._flush(mean=False)
        return 0 if self._sum is None else self._sum[1].long()

    # ---- Stage B ------------------------------------------------------------------------------------
    def aggregate_2d_features_ray_marching(self, projections, features, tsdf):
        """projections [V,B,3,4] un-scaled, features [V,B,C,H',W'], tsdf [B,1,nx,ny,nz] (rm.py:260-307).
        Fills self.points_detection with one [M,3+C] tensor per batch element."""
        self.points_detection = F.rma_points(projections, features, tsdf, self.voxel_dim, self.voxel_size,
                                             self.origin, self.backbone2d_stride, grids=300,
                                             mode=self.ray_marching_type, threshold=self.neus_threshold,
                                             depth_points=self.depth_points, normalize=True)
        for b, pts in enumerate(self.points_detection):
            if pts.shape[0] == 0:      # the reference fails at rm.py:300 when every view was dropped
                raise RuntimeError(f"no valid points for batch element {b}")

    def ray_projection_neus(self, projection, features, tsdf, grids=300, weight_threshold=None):
        """One view; projection [B,3,4] already stride-scaled (rm.py:687-807).  List of [M,4+C] or None."""
        return F.ray_projection(projection, features, tsdf, self.voxel_dim, self.voxel_size, self.origin, grids=grids,
                                mode="neus", threshold=weight_threshold)

    def ray_projection_depth(self, projection, features, tsdf, grids=300, select_grids=None):
        """One view (rm.py:809-956).  List of [M,4+C] or None."""
        return F.ray_projection(projection, features, tsdf, self.voxel_dim, self.voxel_size, self.origin, grids=grids,
                                mode="depth", depth_points=select_grids)


    # ---- hand-off -----------------------------------------------------------------------------------
    def switch_pointcloud(self, points, gt_bboxes, offsets, test):
        """rm.py:339-407: `coord + offset`, sample_points' mask (numpy RNG on the host, as in the reference), ordered
        selection of whole rows on the device, then the caller's augmentation.  Returns (coords, features, boxes)."""
        coords, feats = F.switch_pointcloud(points, offsets, max_points=self.max_points)
        new_boxes = []
        for b in range(len(points)):
            if self.feature_transform is not None and not test:
                coords[b], box = self.feature_transform(coords[b], gt_bboxes[b])
            else:
                box = gt_bboxes[b]
            new_boxes.append(box)
        return coords, feats, new_boxes


def make_detector_class(ray_marching_cls):
    """mmdet adapter: `RayMarchingB200 = make_detector_class(RayMarching)` gives a detector whose five
    aggregation methods and three state attributes come from RayMarchingAggregator while everything else
    (networks, losses, data conversion) stays the reference's.  Register it under its own type name; see
    INTEGRATION.md.  tests/test_adapter_reference_flow.py runs the reference's unmodified forward_train / forward_test on
    such a class (stub networks; CPU stand-ins for the kernels, since the reference tree and a GPU are never on one box)
    and compares with the reference class itself."""

    class RayMarchingB200(ray_marching_cls):
        initialize_volume = RayMarchingAggregator.initialize_volume
        aggregate_2d_features = RayMarchingAggregator.aggregate_2d_features
        clear_3d_features = RayMarchingAggregator.clear_3d_features
        _flush = RayMarchingAggregator._flush
        aggregate_2d_features_ray_marching = RayMarchingAggregator.aggregate_2d_features_ray_marching
        ray_projection_neus = RayMarchingAggregator.ray_projection_neus
        ray_projection_depth = RayMarchingAggregator.ray_projection_depth
        switch_pointcloud = RayMarchingAggregator.switch_pointcloud
        # `volume` / `valid` are assigned by the parent's initialize_volume; the properties take over
        volume = property(RayMarchingAggregator.volume.fget, lambda self, v: None)
        valid = property(RayMarchingAggregator.valid.fget, lambda self, v: None)

    return RayMarchingB200


def make_atlas_class(atlas_cls):
    """The same adapter for the reconstruction-only model `Atlas` (projects/mvsdetection/models/atlas.py:71-...,
    "at.py"), whose `inference1` / `inference2` carry a copy of Stage A (at.py:20-67 `backproject`, :118-179):
    `AtlasB200 = make_atlas_class(Atlas)`.  `inference1` records its (projection, feature) pair (running the 2D
    backbone first when an image is passed, at.py:141-143); `inference2` runs the fused Stage A kernel over all recorded
    views, mean and NaN scrub included (at.py:167-172), then the 3D networks as in the reference (at.py:174-176)."""

    class AtlasB200(atlas_cls):
        initialize_volume = RayMarchingAggregator.initialize_volume
        _flush = RayMarchingAggregator._flush
        clear_3d_features = RayMarchingAggregator.clear_3d_features
        volume = property(RayMarchingAggregator.volume.fget, lambda self, v: None)
        valid = property(RayMarchingAggregator.valid.fget, lambda self, v: None)

        def inference1(self, projection, image=None, feature=None):
            assert ((image is not None and feature is None) or (image is None and feature is not None))
            if feature is None:
                feature = self.backbone2d(self.normalizer(image))
            RayMarchingAggregator.aggregate_2d_features(self, projection, feature)

        def inference2(self, targets=None):
            self.clear_3d_features()
            x = self.backbone3d(self.volume)
            return self.tsdf_head(x, targets)

    return AtlasB200
