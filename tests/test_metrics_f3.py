"""Evaluation metrics around the hot path: ``cd_unit_sphere`` (metrics/metrics.py:176-195) and the P2F
point <-> mesh distances (metrics/metrics.py:196-225 -> metrics/p2m.py:307-375 -> pytorch3d ``_C.point_face_dist_forward``).
pytorch3d is un-vendored: PARITY UNPINNED there; the oracle restates the published algorithm and is pinned on analytic cases,
the CUDA kernel is pinned on the oracle."""
import numpy as np
import pytest
import torch


def _icosphere(sub=2):
    """Unit icosphere (verts [V,3], faces [T,3])."""
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, float) / np.linalg.norm(p) for p in v]
    for _ in range(sub):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = (v[a] + v[b]) / 2
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return torch.tensor(np.array(v), dtype=torch.float32), torch.tensor(f, dtype=torch.int64)


def test_oracle_point_triangle_distance_known_answers():
    """Unit right triangle in z = 0: interior projection -> height^2; beyond a vertex / an edge -> vertex / edge distance;
    a triangle below min_triangle_area is treated as its edges (metrics/p2m.py:20: 5e-3)."""
    from oracle import ops as OO

    tri = torch.tensor([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=torch.float32)
    pts = torch.tensor([[0.2, 0.2, 0.3], [-1, -1, 0], [2, 0, 0], [0.6, 0.6, 0.0], [0.25, 0.25, 0.0]], dtype=torch.float32)
    pd, fd = OO.point_face_dist(pts, tri)
    np.testing.assert_allclose(pd.numpy(), [0.09, 2.0, 1.0, 0.02, 0.0], atol=1e-6)
    assert abs(float(fd[0])) < 1e-7
    small = tri * 0.05                                             # area 1.25e-3 < 5e-3: never "inside"
    p = torch.tensor([[0.0125, 0.0125, 0.3]])
    pd_small, _ = OO.point_face_dist(p, small)
    pd_plane, _ = OO.point_face_dist(p, small, min_triangle_area=0.0)
    assert abs(float(pd_plane[0]) - 0.09) < 1e-6 and float(pd_small[0]) > 0.09 + 1e-5     # edge distance instead of plane distance


@pytest.mark.gpu
def test_point_face_dist_kernel_matches_oracle_and_sphere_geometry():
    from oracle import ops as OO
    from p2pb_b200 import ops

    verts, faces = _icosphere(3)                                   # 1280 faces, area ~ 9.8e-3 each (> 5e-3)
    g = torch.Generator().manual_seed(0)
    pcl = torch.randn(5000, 3, generator=g)
    pcl = pcl / pcl.norm(dim=1, keepdim=True) * (1 + 0.02 * torch.randn(5000, 1, generator=g))
    tris = verts[faces]
    pd_o, fd_o = OO.point_face_dist(pcl, tris)
    p2f, f2p, pd, fd = ops.point_face_dist(pcl.cuda(), verts.cuda(), faces.cuda(), normalize=False)
    np.testing.assert_allclose(pd.cpu().numpy(), pd_o.numpy(), rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(fd.cpu().numpy(), fd_o.numpy(), rtol=2e-4, atol=1e-9)
    assert abs(p2f - float(pd_o.mean())) < 1e-7 and abs(f2p - float(fd_o.mean())) < 1e-7
    assert 2e-4 < p2f < 6e-4                                       # radial noise sigma 0.02 -> E d^2 ~ 4e-4
    # tiny triangles (below min_triangle_area) and the normalised variant run through the same entry point
    v4, f4 = _icosphere(4)
    a, b, _, _ = ops.point_face_dist(pcl.cuda() * 3 + 1, v4.cuda() * 3 + 1, f4.cuda(), normalize=True)
    pd4, fd4 = OO.point_face_dist(pcl, v4[f4])
    assert abs(a - float(pd4.mean())) < 2e-6 and abs(b - float(fd4.mean())) < 2e-6


@pytest.mark.gpu
def test_cd_unit_sphere_matches_restatement():
    from oracle import ops as OO
    from p2pb_b200 import ops

    g = torch.Generator().manual_seed(1)
    ref = torch.randn(2, 3000, 3, generator=g) * torch.tensor([2.0, 1.0, 0.5]) + torch.tensor([5.0, -3.0, 1.0])
    gen = ref[:, :2500] + 0.01 * torch.randn(2, 2500, 3, generator=g)
    c1, c2 = ops.cd_unit_sphere(gen.cuda(), ref.cuda())
    # metrics/metrics.py:139-158, 176-195 in numpy + the oracle's chamfer
    r = ref.numpy().astype(np.float32)
    center = (r.max(1, keepdims=True) + r.min(1, keepdims=True)) / 2
    rc = r - center
    scale = np.sqrt((rc ** 2).sum(-1, keepdims=True)).max(1, keepdims=True)
    d1, d2, _, _ = OO.chamfer_forward(torch.from_numpy((gen.numpy() - center) / scale), torch.from_numpy(rc / scale))
    assert abs(c1 - float(d1.mean())) < 1e-8 and abs(c2 - float(d2.mean())) < 1e-8
    n1, n2 = ops.cd_unit_sphere(gen.cuda(), ref.cuda(), normalize=False)
    assert n1 > c1 and n2 > c2                                     # un-normalised clouds are larger than the unit sphere
