"""Drop-in for the reference's pybind module ``pointnet2_batch_cuda`` (alias ``pointnet2_cuda``,
``third_party/openpoints/cpp/pointnet2_batch/__init__.py:1``): the exact function names and argument order of
``third_party/openpoints/cpp/pointnet2_batch/src/pointnet2_api.cpp:31-47`` for the ops the P2P-Bridge hot path
calls, backed by ``libp2pb_b200.so``.  The reference's unmodified ``models/pvcnn.py`` runs on these.

Differences by design: kernels use torch's current stream; bad input raises ``P2PBError`` (never ``exit()``);
the five ``*_backward`` names exist but raise -- the path is inference only (``p2pb.py:264,337`` run under
``torch.no_grad``).  The legacy PointNet++ exports of the reference module (``*_wrapper``) are not provided:
nothing on the hot path calls them.
"""
from __future__ import annotations

import torch

from . import ops as _ops
from ._lib import P2PBError


def avg_voxelize_forward(features, coords, resolution):
    """vox.cpp:17-44 -> [out [B,C,r^3], ind [B,N], cnt [B,r^3]]"""
    out, ind, cnt = _ops.avg_voxelize(features, coords, resolution)
    return [out, ind, cnt]


def trilinear_devoxelize_forward(r, is_training, coords, features):
    """trilinear_devox.cpp:18-59 -> [outs [B,C,N], inds, wgts] (1-element dummies in eval mode, :48-57)"""
    if is_training:
        raise P2PBError("trilinear_devoxelize_forward(is_training=True) is not supported: inference-only drop-in")
    outs = _ops.trilinear_devoxelize(coords, features, r)
    return [outs, torch.zeros(1, dtype=torch.int32, device=outs.device), torch.zeros(1, device=outs.device)]


def ball_query(centers_coords, points_coords, radius, num_neighbors):
    """pvcnn_ball_query.cpp:6-31"""
    return _ops.ball_query(centers_coords, points_coords, radius, num_neighbors)


def grouping_forward(features, indices):
    """pvcnn_grouping.cpp:6-25"""
    return _ops.grouping(features, indices)


def gather_features_forward(features, indices):
    """pvcnn_sampling.cpp:6-24"""
    return _ops.gather_features(features, indices)


def furthest_point_sampling_forward(coords, num_samples):
    """pvcnn_sampling.cpp:45-61"""
    return _ops.furthest_point_sampling(coords, num_samples)


def three_nearest_neighbors_interpolate_forward(points_coords, centers_coords, centers_features):
    """pvcnn_neighbor_interpolate.cpp:6-41 -> [out [B,C,N], idx [B,3,N], w [B,3,N]]"""
    out, idx, w = _ops.three_nn_interpolate(points_coords, centers_coords, centers_features)
    return [out, idx, w]


def _inference_only(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name}: the B200 drop-in covers the inference path only (training is out of scope)")

    f.__name__ = name
    return f


avg_voxelize_backward = _inference_only("avg_voxelize_backward")
trilinear_devoxelize_backward = _inference_only("trilinear_devoxelize_backward")
three_nearest_neighbors_interpolate_backward = _inference_only("three_nearest_neighbors_interpolate_backward")
grouping_backward = _inference_only("grouping_backward")
gather_features_backward = _inference_only("gather_features_backward")
