"""Torch-tensor front end of the point/voxel kernels in ``libp2pb_b200.so``.

Reference-layout ops (channel-first ``[B,C,N]`` fp32 / int32, contiguous CUDA tensors), one function per native
op the reference's hot path calls (``third_party/openpoints/cpp/pointnet2_batch/src/pointnet2_api.cpp:31-47``).
Outputs are fresh torch tensors owned by the caller; kernels are enqueued on torch's CURRENT stream (the
reference mixes current-stream and legacy-default-stream launches).  CUDA only -- no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import P2PBError, call

_f = ctypes.c_float
_vp = ctypes.c_void_p


def _ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _chk(t, dtype, name, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise P2PBError(f"{name} must be a CUDA tensor (no CPU path)")
    if t.dtype != dtype:
        raise P2PBError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise P2PBError(f"{name} must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise P2PBError(f"{name} must have {ndim} dims, got {tuple(t.shape)}")


def furthest_point_sampling(coords: torch.Tensor, num_samples: int, return_centers: bool = False):
    """coords [B,3,N] -> idx int32 [B,M] (start index 0, reference tie-break); optionally the gathered centres."""
    _chk(coords, torch.float32, "coords", 3)
    B, _, N = coords.shape
    M = int(num_samples)
    idx = torch.empty((B, M), dtype=torch.int32, device=coords.device)
    fuse = return_centers and N <= 16384
    centers = torch.empty((B, 3, M), dtype=torch.float32, device=coords.device) if fuse else None
    scratch = torch.empty((B, N), dtype=torch.float32, device=coords.device) if N > 16384 else None
    with torch.cuda.device(coords.device):
        call("p2pb_furthest_point_sampling", _ptr(coords), B, N, M, _ptr(idx), _ptr(centers), _ptr(scratch), _stream())
    if return_centers:
        if centers is None:
            centers = gather_features(coords, idx)
        return idx, centers
    return idx


def gather_features(features: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    _chk(features, torch.float32, "features", 3)
    _chk(indices, torch.int32, "indices", 2)
    B, C, N = features.shape
    M = indices.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        call("p2pb_gather_features", _ptr(features), _ptr(indices), _ptr(out), B, C, N, M, _stream())
    return out


def grouping(features: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    _chk(features, torch.float32, "features", 3)
    _chk(indices, torch.int32, "indices", 3)
    B, C, N = features.shape
    _, M, U = indices.shape
    out = torch.empty((B, C, M, U), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        call("p2pb_grouping", _ptr(features), _ptr(indices), _ptr(out), B, C, N, M, U, _stream())
    return out


def ball_query(centers: torch.Tensor, points: torch.Tensor, radius: float, num_neighbors: int) -> torch.Tensor:
    _chk(centers, torch.float32, "centers_coords", 3)
    _chk(points, torch.float32, "points_coords", 3)
    B, _, M = centers.shape
    N = points.shape[2]
    out = torch.empty((B, M, int(num_neighbors)), dtype=torch.int32, device=centers.device)
    with torch.cuda.device(centers.device):
        call("p2pb_ball_query", _ptr(centers), _ptr(points), B, M, N, _f(radius), int(num_neighbors), _ptr(out), _stream())
    return out


def three_nn_interpolate(points: torch.Tensor, centers: torch.Tensor, centers_features: torch.Tensor):
    _chk(points, torch.float32, "points_coords", 3)
    _chk(centers, torch.float32, "centers_coords", 3)
    _chk(centers_features, torch.float32, "centers_features", 3)
    B, _, N = points.shape
    M = centers.shape[2]
    C = centers_features.shape[1]
    dev = points.device
    out = torch.empty((B, C, N), dtype=torch.float32, device=dev)
    idx = torch.empty((B, 3, N), dtype=torch.int32, device=dev)
    w = torch.empty((B, 3, N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("p2pb_three_nn_interpolate", _ptr(points), _ptr(centers), _ptr(centers_features), B, C, N, M,
             _ptr(out), _ptr(idx), _ptr(w), _stream())
    return out, idx, w


def avg_voxelize(features: torch.Tensor, coords: torch.Tensor, resolution: int):
    """features [B,C,N], int32 voxel coords [B,3,N] -> (grid [B,C,r^3], ind [B,N], cnt [B,r^3])."""
    _chk(features, torch.float32, "features", 3)
    _chk(coords, torch.int32, "coords", 3)
    B, C, N = features.shape
    r = int(resolution)
    dev = features.device
    out = torch.empty((B, C, r ** 3), dtype=torch.float32, device=dev)
    ind = torch.empty((B, N), dtype=torch.int32, device=dev)
    cnt = torch.empty((B, r ** 3), dtype=torch.int32, device=dev)
    order = torch.empty((B, N), dtype=torch.int32, device=dev)
    start = torch.empty((B, r ** 3), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        call("p2pb_avg_voxelize", _ptr(features), _ptr(coords), B, C, N, r, _ptr(out), _ptr(ind), _ptr(cnt),
             _ptr(order), _ptr(start), _stream())
    return out, ind, cnt


def voxel_prep(coords: torch.Tensor, resolution: int, normalize: bool = True, eps: float = 0.0):
    """Fused ``Voxelization.forward`` coordinate prep (models/pvcnn.py:215-231) + CSR build.

    coords [B,3,N] -> dict(norm_coords [B,3,N], ind [B,N], order [B,N], start [B,r^3], cnt [B,r^3])."""
    _chk(coords, torch.float32, "coords", 3)
    B, _, N = coords.shape
    r = int(resolution)
    dev = coords.device
    nc = torch.empty_like(coords)
    ind = torch.empty((B, N), dtype=torch.int32, device=dev)
    order = torch.empty((B, N), dtype=torch.int32, device=dev)
    start = torch.empty((B, r ** 3), dtype=torch.int32, device=dev)
    cnt = torch.empty((B, r ** 3), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        call("p2pb_voxel_prep", _ptr(coords), B, N, r, int(bool(normalize)), _f(eps), _ptr(nc), _ptr(ind), _ptr(order),
             _ptr(start), _ptr(cnt), _stream())
    return {"norm_coords": nc, "ind": ind, "order": order, "start": start, "cnt": cnt, "r": r}


def trilinear_devoxelize(coords: torch.Tensor, grid: torch.Tensor, resolution: int) -> torch.Tensor:
    """coords [B,3,N] in [0,r-1], grid [B,C,r^3] -> [B,C,N] (inference branch of the reference op)."""
    _chk(coords, torch.float32, "coords", 3)
    _chk(grid, torch.float32, "features", 3)
    B, C, _ = grid.shape
    N = coords.shape[2]
    out = torch.empty((B, C, N), dtype=torch.float32, device=grid.device)
    with torch.cuda.device(grid.device):
        call("p2pb_trilinear_devoxelize", _ptr(coords), _ptr(grid), B, C, N, int(resolution), _ptr(out), _stream())
    return out


def nm_distance(xyz1: torch.Tensor, xyz2: torch.Tensor):
    """xyz1 [B,n,3], xyz2 [B,m,3] -> (squared NN distance [B,n], index int32 [B,n]); chamfer3D.cu:12-134."""
    _chk(xyz1, torch.float32, "xyz1", 3)
    _chk(xyz2, torch.float32, "xyz2", 3)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    dist = torch.empty((B, n), dtype=torch.float32, device=dev)
    idx = torch.empty((B, n), dtype=torch.int32, device=dev)
    scratch = torch.empty((B, n), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        call("p2pb_nm_distance", _ptr(xyz1), _ptr(xyz2), B, n, m, _ptr(dist), _ptr(idx), _ptr(scratch), _stream())
    return dist, idx


def chamfer_forward(xyz1: torch.Tensor, xyz2: torch.Tensor):
    """``chamfer_3D.forward`` results: dist1, dist2, idx1, idx2 (metrics/chamfer3D/chamfer_cuda.cpp)."""
    d1, i1 = nm_distance(xyz1, xyz2)
    d2, i2 = nm_distance(xyz2, xyz1)
    return d1, d2, i1, i2


def calculate_cd(pred: torch.Tensor, gt: torch.Tensor):
    """metrics/metrics.py:56-83 ``calculate_cd_cuda``: per-sample mean(d1)+mean(d2) of squared NN distances."""
    if pred.shape[-1] != 3:
        pred = pred.transpose(-1, -2)
        gt = gt.transpose(-1, -2)
    d1, d2, _, _ = chamfer_forward(pred.contiguous(), gt.contiguous())
    return (d1.mean(dim=1) + d2.mean(dim=1)).cpu().tolist()


def emd_approx(xyz1: torch.Tensor, xyz2: torch.Tensor) -> torch.Tensor:
    """``emd_cuda.matchcost_forward(xyz1, xyz2, emd_cuda.approxmatch_forward(xyz1, xyz2))`` without the match matrix:
    xyz1 [B,n,3], xyz2 [B,m,3] -> un-normalised cost [B] (metrics/PyTorchEMD/cuda/emd_kernel.cu:33-165, 211-253)."""
    _chk(xyz1, torch.float32, "xyz1", 3)
    _chk(xyz2, torch.float32, "xyz2", 3)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    cost = torch.empty((B,), dtype=torch.float32, device=dev)
    scratch = torch.empty((B * (3 * n + 2 * m),), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("p2pb_emd_approx", _ptr(xyz1), _ptr(xyz2), B, n, m, _ptr(cost), _ptr(scratch), _stream())
    return cost


def earth_mover_distance_nograd(xyz1: torch.Tensor, xyz2: torch.Tensor, transpose: bool = True) -> torch.Tensor:
    """metrics/PyTorchEMD/emd_nograd.py:19-44: approximate EMD / N; inputs [B,3,N] (transpose=True) or [B,N,3]."""
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1 = xyz1.transpose(1, 2)
        xyz2 = xyz2.transpose(1, 2)
    assert xyz1.shape[-1] == 3, f"require it to be B,N,3; get: {xyz1.shape}"
    return emd_approx(xyz1.contiguous(), xyz2.contiguous()) / float(xyz1.shape[1])


def knn_points(queries: torch.Tensor, points: torch.Tensor, K: int, return_dist: bool = False):
    """K nearest ``points [N,3]`` of every ``queries [Q,3]`` row, ascending squared distance, ties by lower index
    (``pytorch3d.ops.knn_points(..., return_sorted=True)`` as used at denoise_object.py:90-91) -> idx int32 [Q,K]."""
    _chk(queries, torch.float32, "queries", 2)
    _chk(points, torch.float32, "points", 2)
    Q, N = queries.shape[0], points.shape[0]
    dev = points.device
    idx = torch.empty((Q, K), dtype=torch.int32, device=dev)
    dist = torch.empty((Q, K), dtype=torch.float32, device=dev) if return_dist else None
    with torch.cuda.device(dev):
        call("p2pb_knn_points", _ptr(queries), _ptr(points), Q, N, int(K), _ptr(idx), _ptr(dist), _stream())
    return (idx, dist) if return_dist else idx


def radius_count(centers: torch.Tensor, points: torch.Tensor, radius: float) -> torch.Tensor:
    """Number of ``points [N,3]`` within ``radius`` of each ``centers [P,3]`` row -> int32 [P] (first pass of radius_query)."""
    _chk(centers, torch.float32, "centers", 2)
    _chk(points, torch.float32, "points", 2)
    counts = torch.empty((centers.shape[0],), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        call("p2pb_radius_count", _ptr(centers), _ptr(points), centers.shape[0], points.shape[0], _f(float(radius)), _ptr(counts), _stream())
    return counts


def radius_query(centers: torch.Tensor, points: torch.Tensor, radius: float):
    """All ``points [N,3]`` within ``radius`` of each ``centers [P,3]`` row (``KDTree.query_radius`` of
    denoise_room.py:454-465) -> (offsets int64 [P+1], indices int32 [offsets[-1]]), indices ascending per centre."""
    _chk(centers, torch.float32, "centers", 2)
    _chk(points, torch.float32, "points", 2)
    P, N = centers.shape[0], points.shape[0]
    dev = points.device
    counts = torch.empty((P,), dtype=torch.int32, device=dev)
    offsets = torch.zeros((P + 1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        call("p2pb_radius_count", _ptr(centers), _ptr(points), P, N, _f(float(radius)), _ptr(counts), _stream())
        offsets[1:] = torch.cumsum(counts, 0)
        indices = torch.empty((int(offsets[-1].item()),), dtype=torch.int32, device=dev)
        call("p2pb_radius_fill", _ptr(centers), _ptr(points), P, N, _f(float(radius)), _ptr(offsets), _ptr(indices), _stream())
    return offsets, indices



# ---- room sweep: device-side patch creation and reassembly (csrc/room.cu; denoise_room.py:352-421, 141-146, 262-289) ---------
def room_pad_patches(room: torch.Tensor, off: torch.Tensor, csr: torch.Tensor, job_patch: torch.Tensor, M: int, seed: int,
                     pre=None, job_key: torch.Tensor = None):
    """Under-full radius patches (n < M) -> (xyz [J,M,3], idx int32 [J,M], cut int32 [J]).  ``pre`` = (pre_off int64 [J+1],
    pre_idx int32, pre_noise fp32 [.,3]) replaces the counter-based RNG by host-drawn randoms (``--strict_ref``); ``job_key``
    int32 [J] = the numbers the counter RNG is keyed by (global patch numbers; default ``job_patch``)."""
    _chk(room, torch.float32, "room", 2)
    _chk(off, torch.int64, "off", 1)
    _chk(csr, torch.int32, "csr", 1)
    _chk(job_patch, torch.int32, "job_patch", 1)
    J, dev = job_patch.shape[0], room.device
    xyz = torch.empty((J, M, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((J, M), dtype=torch.int32, device=dev)
    cut = torch.empty((J,), dtype=torch.int32, device=dev)
    po, pi, pn = (None, None, None) if pre is None else pre
    with torch.cuda.device(dev):
        call("p2pb_room_pad_patches", _ptr(room), _ptr(off), _ptr(csr), _ptr(job_patch), _ptr(job_key), J, int(M), ctypes.c_ulonglong(int(seed) & (2 ** 64 - 1)),
             _ptr(po), _ptr(pi), _ptr(pn), _ptr(xyz), _ptr(idx), _ptr(cut), _stream())
    return xyz, idx, cut


def room_fps_patches(room: torch.Tensor, off: torch.Tensor, csr: torch.Tensor, job_patch: torch.Tensor, job_start: torch.Tensor,
                     n_max: int, M: int):
    """Over-full radius patches (n >= M): exact FPS of M points from local start index ``job_start`` per (patch, replica) job
    -> (xyz [J,M,3] in FPS order, idx int32 [J,M])."""
    _chk(room, torch.float32, "room", 2)
    _chk(off, torch.int64, "off", 1)
    _chk(csr, torch.int32, "csr", 1)
    _chk(job_patch, torch.int32, "job_patch", 1)
    _chk(job_start, torch.int32, "job_start", 1)
    J, dev = job_patch.shape[0], room.device
    xyz = torch.empty((J, M, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((J, M), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        call("p2pb_room_fps_patches", _ptr(room), _ptr(off), _ptr(csr), _ptr(job_patch), _ptr(job_start), J, int(n_max), int(M),
             _ptr(xyz), _ptr(idx), _stream())
    return xyz, idx


def patch_normalize(xyz: torch.Tensor):
    """xyz [P,M,3] -> (x_start [P,3,M] fp32 centred / max-norm scaled, center f64 [P,3], scale f64 [P])."""
    _chk(xyz, torch.float32, "xyz", 3)
    P, M, _ = xyz.shape
    dev = xyz.device
    x = torch.empty((P, 3, M), dtype=torch.float32, device=dev)
    c = torch.empty((P, 3), dtype=torch.float64, device=dev)
    s = torch.empty((P,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        call("p2pb_patch_normalize", _ptr(xyz), P, M, _ptr(x), _ptr(c), _ptr(s), _stream())
    return x, c, s


def room_accumulate(x_pred: torch.Tensor, center: torch.Tensor, scale: torch.Tensor, idx: torch.Tensor, cut: torch.Tensor,
                    sum_fixed: torch.Tensor, count: torch.Tensor) -> None:
    """Add the de-normalised predictions of P patches into ``sum_fixed int64 [N,3]`` (2^-40 units) / ``count int32 [N]``."""
    _chk(x_pred, torch.float32, "x_pred", 3)
    _chk(center, torch.float64, "center", 2)
    _chk(scale, torch.float64, "scale", 1)
    _chk(idx, torch.int32, "idx", 2)
    _chk(cut, torch.int32, "cut", 1)
    _chk(sum_fixed, torch.int64, "sum_fixed", 2)
    _chk(count, torch.int32, "count", 1)
    P, _, M = x_pred.shape
    with torch.cuda.device(x_pred.device):
        call("p2pb_room_accumulate", _ptr(x_pred), _ptr(center), _ptr(scale), _ptr(idx), _ptr(cut), P, M, _ptr(sum_fixed), _ptr(count),
             _stream())


# ---- evaluation metrics at scale (metrics/metrics.py:139-225, metrics/p2m.py:307-375) --------------------------------------
def normalize_sphere(pc: torch.Tensor, radius: float = 1.0):
    """metrics/metrics.py:139-158: bounding-box centre, max-norm scale -> (pc, center [B,1,3], scale [B,1,1])."""
    p_max = pc.max(dim=-2, keepdim=True)[0]
    p_min = pc.min(dim=-2, keepdim=True)[0]
    center = (p_max + p_min) / 2
    pc = pc - center
    scale = (pc ** 2).sum(dim=-1, keepdim=True).sqrt().max(dim=-2, keepdim=True)[0] / radius
    return pc / scale, center, scale


def cd_unit_sphere(gen: torch.Tensor, ref: torch.Tensor, normalize: bool = True):
    """metrics/metrics.py:176-195: Chamfer distance of ``gen [B,N,3]`` vs ``ref [B,M,3]`` after normalising BOTH with the
    reference cloud's unit-sphere transform -> (mean squared distance gen->ref, ref->gen) as floats."""
    if normalize:
        ref, center, scale = normalize_sphere(ref)
        gen = (gen - center) / scale
    d1, d2, _, _ = chamfer_forward(gen.contiguous().float(), ref.contiguous().float())
    return d1.mean().item(), d2.mean().item()


def point_face_dist(pcl: torch.Tensor, verts: torch.Tensor, faces: torch.Tensor, normalize: bool = True,
                    min_triangle_area: float = 5e-3):
    """metrics/metrics.py:196-225 ``point_face_dist`` (P2F metric): ``pcl [P,3]``, mesh ``verts [V,3]`` / ``faces int [T,3]``,
    both normalised with the MESH's unit-sphere transform -> (mean squared point->face distance, mean squared face->point
    distance), the two terms of ``point_mesh_face_distance_custom`` (metrics/p2m.py:307-375; default min_triangle_area 5e-3,
    p2m.py:20)."""
    assert pcl.dim() == 2 and verts.dim() == 2 and faces.dim() == 2, "Batch is not supported."
    if normalize:
        v, center, scale = normalize_sphere(verts.unsqueeze(0))
        verts = v[0]
        pcl = ((pcl.unsqueeze(0) - center) / scale)[0]
    pcl = pcl.contiguous().float()
    tris = verts.float()[faces.long()].contiguous()          # [T,3,3]
    _chk(pcl, torch.float32, "pcl", 2)
    P, T = pcl.shape[0], tris.shape[0]
    pd = torch.empty((P,), dtype=torch.float32, device=pcl.device)
    fd = torch.empty((T,), dtype=torch.float32, device=pcl.device)
    with torch.cuda.device(pcl.device):
        call("p2pb_point_face_dist", _ptr(pcl), P, _ptr(tris), T, _f(float(min_triangle_area)), _ptr(pd), _ptr(fd), _stream())
    return pd.mean().item(), fd.mean().item(), pd, fd
