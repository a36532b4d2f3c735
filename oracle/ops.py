"""ctypes front-end of ``oracle/p2pb_oracle.c`` -- TEST INFRASTRUCTURE ONLY (see the header of that file).

Exposes the reference's pybind names (``pointnet2_api.cpp:31-47``) over CPU torch tensors so that the
oracle model (``oracle/model.py``) and the reference's real Python model code (imported from
``/root/reference`` by ``oracle/gen_golden.py``) can both run on the host.

Nothing under ``p2pb_b200/`` may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "p2pb_oracle.c")
_SO = os.path.join(_HERE, "libp2pb_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off -fopenmp: contraction OFF so only the explicit fmaf() chains contract."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
               "-o", _SO, _SRC, "-lm"]
        subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def voxel_coords(coords, r, normalize=True, eps=0.0):
    """models/pvcnn.py:215-231 coordinate prep -> (norm_coords f32 [B,3,N], vox_coords i32 [B,3,N])."""
    coords = coords.contiguous().float()
    B, _, N = coords.shape
    nc = torch.empty_like(coords)
    vc = torch.empty((B, 3, N), dtype=torch.int32)
    lib().ora_voxel_coords(_f(coords), B, N, int(r), int(bool(normalize)), ctypes.c_float(eps), _f(nc), _i(vc))
    return nc, vc


def avg_voxelize_forward(features, coords, resolution):
    features = features.contiguous().float()
    coords = coords.contiguous().int()
    B, C, N = features.shape
    r = int(resolution)
    out = torch.empty((B, C, r ** 3), dtype=torch.float32)
    ind = torch.empty((B, N), dtype=torch.int32)
    cnt = torch.empty((B, r ** 3), dtype=torch.int32)
    lib().ora_avg_voxelize_forward(_f(features), _i(coords), B, C, N, r, _f(out), _i(ind), _i(cnt))
    return [out, ind, cnt]


def trilinear_devoxelize_forward(r, is_training, coords, features):
    assert not is_training, "oracle restates the inference path only (trilinear_devox.cpp:48-57)"
    coords = coords.contiguous().float()
    features = features.contiguous().float()
    B, C = features.shape[:2]
    N = coords.shape[2]
    outs = torch.empty((B, C, N), dtype=torch.float32)
    lib().ora_trilinear_devoxelize_forward(_f(coords), _f(features), B, C, N, int(r), _f(outs))
    return [outs, torch.zeros(1, dtype=torch.int32), torch.zeros(1)]


def ball_query(centers_coords, points_coords, radius, num_neighbors):
    centers_coords = centers_coords.contiguous().float()
    points_coords = points_coords.contiguous().float()
    B, _, M = centers_coords.shape
    N = points_coords.shape[2]
    idx = torch.empty((B, M, num_neighbors), dtype=torch.int32)
    lib().ora_ball_query(_f(centers_coords), _f(points_coords), B, M, N, ctypes.c_float(radius), int(num_neighbors), _i(idx))
    return idx


def grouping_forward(features, indices):
    features = features.contiguous().float()
    indices = indices.contiguous().int()
    B, C, N = features.shape
    _, M, U = indices.shape
    out = torch.empty((B, C, M, U), dtype=torch.float32)
    lib().ora_grouping_forward(_f(features), _i(indices), B, C, N, M, U, _f(out))
    return out


def gather_features_forward(features, indices):
    features = features.contiguous().float()
    indices = indices.contiguous().int()
    B, C, N = features.shape
    M = indices.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32)
    lib().ora_gather_features_forward(_f(features), _i(indices), B, C, N, M, _f(out))
    return out


def furthest_point_sampling_forward(coords, num_samples):
    coords = coords.contiguous().float()
    B, _, N = coords.shape
    idx = torch.empty((B, num_samples), dtype=torch.int32)
    lib().ora_furthest_point_sampling(_f(coords), B, N, int(num_samples), _i(idx))
    return idx


furthest_point_sampling = furthest_point_sampling_forward  # _pvcnn_backend name (third_party/pvcnn/functional/src/bindings.cpp:15)


def three_nearest_neighbors_interpolate_forward(points_coords, centers_coords, centers_features):
    points_coords = points_coords.contiguous().float()
    centers_coords = centers_coords.contiguous().float()
    centers_features = centers_features.contiguous().float()
    B, _, N = points_coords.shape
    M = centers_coords.shape[2]
    C = centers_features.shape[1]
    out = torch.empty((B, C, N), dtype=torch.float32)
    idx = torch.empty((B, 3, N), dtype=torch.int32)
    w = torch.empty((B, 3, N), dtype=torch.float32)
    lib().ora_three_nn_interpolate_forward(_f(points_coords), _f(centers_coords), _f(centers_features),
                                           B, C, N, M, _f(out), _i(idx), _f(w))
    return [out, idx, w]


def nm_distance(xyz1, xyz2):
    """chamfer3D.cu:12-134: (dist [B,n] squared NN distance, idx [B,n]) of xyz1 [B,n,3] into xyz2 [B,m,3]."""
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d = torch.empty((B, n), dtype=torch.float32)
    i = torch.empty((B, n), dtype=torch.int32)
    lib().ora_nm_distance(_f(xyz1), _f(xyz2), B, n, m, _f(d), _i(i))
    return d, i


def chamfer_forward(xyz1, xyz2):
    """chamfer_cuda.cpp forward: dist1, dist2, idx1, idx2."""
    d1, i1 = nm_distance(xyz1, xyz2)
    d2, i2 = nm_distance(xyz2, xyz1)
    return d1, d2, i1, i2


def calculate_cd(pred, gt):
    """metrics/metrics.py:56-83 calculate_cd_cuda: mean(d1)+mean(d2) of squared NN distances, per sample."""
    if pred.shape[-1] != 3:
        pred = pred.transpose(-1, -2)
        gt = gt.transpose(-1, -2)
    d1, d2, _, _ = chamfer_forward(pred, gt)
    return (d1.mean(dim=1) + d2.mean(dim=1)).tolist()


def emd_approxmatch_cost(xyz1, xyz2):
    """emd_kernel.cu approxmatch+matchcost: cost [B] (un-normalised; emd_nograd.py:43 divides by N)."""
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = torch.empty((B, m, n), dtype=torch.float32)
    cost = torch.empty((B,), dtype=torch.float32)
    lib().ora_emd_approxmatch_cost(_f(xyz1), _f(xyz2), B, n, m, _f(match), _f(cost))
    return cost, match


def knn_points(queries, points, K):
    """pytorch3d.ops.knn_points contract (squared L2, K smallest ascending, ties by index): idx int32 [Q,K], dist [Q,K]."""
    queries = queries.contiguous().float()
    points = points.contiguous().float()
    Q, N = queries.shape[0], points.shape[0]
    idx = torch.empty((Q, K), dtype=torch.int32)
    dist = torch.empty((Q, K), dtype=torch.float32)
    lib().ora_knn_points(_f(queries), _f(points), Q, N, int(K), _i(idx), _f(dist))
    return idx, dist


def radius_query(centers, points, radius):
    """KDTree.query_radius contract restated (ascending indices): (offsets int64 [P+1], indices int32)."""
    centers = centers.contiguous().float()
    points = points.contiguous().float()
    P, N = centers.shape[0], points.shape[0]
    counts = torch.empty((P,), dtype=torch.int32)
    lib().ora_radius_query(_f(centers), _f(points), P, N, ctypes.c_float(radius), _i(counts), None, None)
    offsets = torch.zeros((P + 1,), dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    indices = torch.empty((int(offsets[-1]),), dtype=torch.int32)
    lib().ora_radius_query(_f(centers), _f(points), P, N, ctypes.c_float(radius), _i(counts),
                           ctypes.c_void_p(offsets.data_ptr()), _i(indices))
    return offsets, indices


def point_face_dist(pts, tris, min_triangle_area=5e-3):
    """pytorch3d point_face / face_point squared distances restated (metrics/p2m.py:307-375; PARITY UNPINNED, see p2pb_oracle.c):
    pts [P,3], tris [T,3,3] -> (point_dist [P], face_dist [T])."""
    pts = pts.contiguous().float()
    tris = tris.contiguous().float()
    P, T = pts.shape[0], tris.shape[0]
    pd, fd = torch.empty((P,), dtype=torch.float32), torch.empty((T,), dtype=torch.float32)
    lib().ora_point_face_dist(_f(pts), P, _f(tris), T, ctypes.c_float(min_triangle_area), _f(pd), _f(fd))
    return pd, fd


_BACKWARD = ["avg_voxelize_backward", "trilinear_devoxelize_backward", "three_nearest_neighbors_interpolate_backward",
             "grouping_backward", "gather_features_backward"]


def _no_backward(*a, **k):
    raise NotImplementedError("oracle restates the inference path only")


for _n in _BACKWARD:
    globals()[_n] = _no_backward
