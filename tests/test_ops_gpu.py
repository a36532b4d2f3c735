"""GPU parity tests of the point/voxel kernels, called through the C ABI (ctypes), against the oracle on the same
seeded inputs.  Bar: bit-exact for indices AND for every fp32 output (same contraction order, deterministic sums)."""
import numpy as np
import pytest
import torch

from tests.helpers import cloud

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from oracle import ops as OO
    from p2pb_b200 import ops, pointnet2_batch_cuda as P

    return OO, ops, P


CASES = [(2, 2048, 512, None), (3, 1024, 256, 8.0), (2, 8192, 2048, None), (4, 128, 32, None), (2, 32, 8, 4.0),
         (1, 1000, 250, None)]


@pytest.mark.parametrize("B,N,M,snap", CASES)
def test_fps_gather_ball_group_bit_exact(mods, B, N, M, snap):
    OO, ops, P = mods
    coords = cloud(B, N, seed=N + M, snap=snap)
    cd = coords.cuda()
    idx = P.furthest_point_sampling_forward(cd, M)
    ref_idx = OO.furthest_point_sampling_forward(coords, M)
    assert torch.equal(idx.cpu(), ref_idx)
    idx2, centers = ops.furthest_point_sampling(cd, M, return_centers=True)
    ref_centers = OO.gather_features_forward(coords, ref_idx)
    assert torch.equal(idx2.cpu(), ref_idx) and torch.equal(centers.cpu(), ref_centers)
    assert torch.equal(P.gather_features_forward(cd, idx).cpu(), ref_centers)
    for radius in (0.1, 0.4):
        nidx = P.ball_query(centers, cd, radius, 32)
        ref_n = OO.ball_query(ref_centers, coords, radius, 32)
        assert torch.equal(nidx.cpu(), ref_n), radius
    feats = torch.randn(B, 7, N, generator=torch.Generator().manual_seed(1))
    g = P.grouping_forward(feats.cuda(), nidx)
    assert torch.equal(g.cpu(), OO.grouping_forward(feats, ref_n))


# (1, 1000, 250, 3): M not a multiple of 4 (scalar tail of the 16-byte centre loads); (40, 4096, 64, 2): more point blocks per patch than
# the launcher's residency cap allows CTAs -> the grid-stride loop over point blocks
@pytest.mark.parametrize("B,N,M,C", [(2, 2048, 512, 16), (3, 512, 128, 5), (2, 32, 8, 64), (1, 8192, 2048, 4), (1, 1000, 250, 3),
                                     (40, 4096, 64, 2)])
def test_three_nn_interpolate_bit_exact(mods, B, N, M, C):
    OO, ops, P = mods
    pts = cloud(B, N, seed=3)
    ctr = pts[:, :, ::N // M].contiguous()  # centres coincide with points -> zero distances hit the 1e-10 clamp
    cf = torch.randn(B, C, M, generator=torch.Generator().manual_seed(2))
    out, idx, w = P.three_nearest_neighbors_interpolate_forward(pts.cuda(), ctr.cuda(), cf.cuda())
    ro, ri, rw = OO.three_nearest_neighbors_interpolate_forward(pts, ctr, cf)
    assert torch.equal(idx.cpu(), ri)
    assert torch.equal(w.cpu(), rw)
    assert torch.equal(out.cpu(), ro)


@pytest.mark.parametrize("B,N,C,r", [(2, 2048, 35, 32), (3, 512, 8, 16), (2, 128, 16, 8), (1, 8192, 4, 32), (2, 100, 3, 4)])
def test_voxelize_devoxelize_bit_exact(mods, B, N, C, r):
    OO, ops, P = mods
    coords = cloud(B, N, seed=7 + r)
    feats = torch.randn(B, C, N, generator=torch.Generator().manual_seed(4))
    prep = ops.voxel_prep(coords.cuda(), r)
    nc, vc = OO.voxel_coords(coords, r)
    assert torch.equal(prep["norm_coords"].cpu(), nc)
    ref_out, ref_ind, ref_cnt = OO.avg_voxelize_forward(feats, vc, r)
    assert torch.equal(prep["ind"].cpu(), ref_ind) and torch.equal(prep["cnt"].cpu(), ref_cnt)
    # CSR: order is sorted by (voxel, point index) and start/cnt partition it
    order, start, cnt = prep["order"].cpu(), prep["start"].cpu(), prep["cnt"].cpu()
    for b in range(B):
        ind_sorted = ref_ind[b][order[b].long()]
        assert torch.all(ind_sorted[1:] >= ind_sorted[:-1])
        assert sorted(order[b].tolist()) == list(range(N))
        occ = torch.nonzero(cnt[b]).flatten()
        assert int(cnt[b].sum()) == N
        v = occ[len(occ) // 2].item()
        seg = order[b][start[b][v]: start[b][v] + cnt[b][v]]
        assert torch.all(ref_ind[b][seg.long()] == v) and torch.all(seg[1:] > seg[:-1])
    out, ind, cnt2 = P.avg_voxelize_forward(feats.cuda(), vc.cuda(), r)
    assert torch.equal(ind.cpu(), ref_ind) and torch.equal(cnt2.cpu(), ref_cnt)
    assert torch.equal(out.cpu(), ref_out)  # deterministic index-ordered sums: bit-exact, unlike the reference's atomics
    grid = torch.randn(B, C, r ** 3, generator=torch.Generator().manual_seed(5))
    dv = P.trilinear_devoxelize_forward(r, False, prep["norm_coords"], grid.cuda())[0]
    assert torch.equal(dv.cpu(), OO.trilinear_devoxelize_forward(r, False, nc, grid)[0])


def test_voxelize_all_points_in_one_voxel_and_r_minus_1_corner(mods):
    OO, ops, P = mods
    B, N, C, r = 1, 64, 4, 8
    vc = torch.zeros(B, 3, N, dtype=torch.int32)
    vc[:, :, N // 2:] = r - 1
    feats = torch.randn(B, C, N, generator=torch.Generator().manual_seed(9))
    out, ind, cnt = P.avg_voxelize_forward(feats.cuda(), vc.cuda(), r)
    ro, ri, rc = OO.avg_voxelize_forward(feats, vc, r)
    assert torch.equal(out.cpu(), ro) and torch.equal(cnt.cpu(), rc)
    # devox exactly at integer coordinates incl. r-1: hi-corner must not be touched (trilinear_devox_gpu.cu:66-68)
    coords = torch.tensor([[[0.0, float(r - 1), 3.0], [0.0, float(r - 1), 2.5], [0.0, float(r - 1), 7.0]]])
    grid = torch.randn(1, C, r ** 3, generator=torch.Generator().manual_seed(3))
    dv = P.trilinear_devoxelize_forward(r, False, coords.cuda(), grid.cuda())[0]
    assert torch.equal(dv.cpu(), OO.trilinear_devoxelize_forward(r, False, coords, grid)[0])


@pytest.mark.parametrize("B,n,m", [(2, 2048, 2048), (1, 8192, 5000), (3, 100, 1500), (64, 2048, 2048)])
def test_chamfer_bit_exact(mods, B, n, m):
    OO, ops, P = mods
    a = cloud(B, n, seed=1, snap=16.0 if B == 3 else None).transpose(1, 2).contiguous()
    b = (cloud(B, m, seed=2, snap=16.0 if B == 3 else None) * 1.02).transpose(1, 2).contiguous()
    if B == 64:  # full bench batch: property checks only (self-distance is zero, symmetric use)
        d1, d2, i1, i2 = ops.chamfer_forward(a.cuda(), a.cuda())
        assert float(d1.abs().max()) == 0.0 and torch.equal(i1.cpu(), torch.arange(n, dtype=torch.int32).expand(B, n))
        return
    d1, d2, i1, i2 = ops.chamfer_forward(a.cuda(), b.cuda())
    r1, r2, j1, j2 = OO.chamfer_forward(a, b)
    assert torch.equal(d1.cpu(), r1) and torch.equal(d2.cpu(), r2)
    assert torch.equal(i1.cpu(), j1) and torch.equal(i2.cpu(), j2)
    assert ops.calculate_cd(a.cuda(), b.cuda()) == pytest.approx(OO.calculate_cd(a, b), rel=1e-6)


def test_large_cloud_fps_global_path(mods):
    """N > 16384 takes the global-memory FPS kernel (object-level seed/merge FPS); same selection rule."""
    OO, ops, P = mods
    coords = cloud(1, 20000, seed=5)
    idx = P.furthest_point_sampling_forward(coords.cuda(), 300)
    assert torch.equal(idx.cpu(), OO.furthest_point_sampling_forward(coords, 300))


@pytest.mark.parametrize("tag", ["a", "tie"])
def test_kernels_match_reference_kernels(mods, ref_ops, tag):
    """Against outputs of the reference's OWN compiled CUDA kernels (tests/golden/ref_ops.npz)."""
    OO, ops, P = mods
    g = lambda k: torch.from_numpy(ref_ops[f"{tag}_{k}"]).cuda()
    coords, feats = g("coords"), g("feats")
    idx = P.furthest_point_sampling_forward(coords, 512)
    assert torch.equal(idx, g("fps_idx"))
    centers = P.gather_features_forward(coords, idx)
    nidx = P.ball_query(centers, coords, 0.1, 32)
    assert torch.equal(nidx, g("ball_idx"))
    assert torch.equal(P.grouping_forward(feats, nidx), g("grouped"))
    out, iidx, iw = P.three_nearest_neighbors_interpolate_forward(coords, centers, g("cfeat"))
    assert torch.equal(iidx, g("interp_idx")) and torch.equal(iw, g("interp_w")) and torch.equal(out, g("interp"))
    vo, vind, vcnt = P.avg_voxelize_forward(feats, g("vox"), 16)
    assert torch.equal(vind, g("vox_ind")) and torch.equal(vcnt, g("vox_cnt"))
    spread = (g("vox_out") - g("vox_out2")).abs().max().item()
    assert (vo - g("vox_out")).abs().max().item() <= max(2e-6, 4 * spread)
    dv = P.trilinear_devoxelize_forward(16, False, g("norm_coords"), g("grid"))[0]
    assert torch.equal(dv, g("devox"))
    d1, i1 = ops.nm_distance(coords.transpose(1, 2).contiguous(), (coords[:, :, :1024] + 0.01).transpose(1, 2).contiguous())
    assert torch.equal(d1, g("cd_d1")) and torch.equal(i1, g("cd_i1"))


def test_emd_matches_oracle_and_reference_kernels(ref_ops):
    """Match-free multi-CTA EMD == the reference's own approxmatch+matchcost kernels (golden from oracle/_ref on a B200),
    the reference's known answer (test_emd_loss.py:6-20 -> 0.71) and the C oracle on ragged sizes (n != m).
    fp tolerance 2e-3 relative: __expf and a different summation order, as between the reference GPU run and the oracle."""
    from oracle import ops as OO
    from p2pb_b200 import ops

    a, b = torch.from_numpy(ref_ops["emd_a"]), torch.from_numpy(ref_ops["emd_b"])
    cost = ops.emd_approx(a.cuda(), b.cuda()).cpu().numpy()
    np.testing.assert_allclose(cost, ref_ops["emd_cost"], rtol=2e-3)
    p1 = torch.tensor([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]]).repeat(3, 1, 1)
    p2 = torch.tensor([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]]).repeat(3, 1, 1)
    np.testing.assert_allclose(ops.emd_approx(p1.cuda(), p2.cuda()).cpu().numpy(), np.full(3, 0.71), rtol=2e-3)
    g = torch.Generator().manual_seed(9)
    for n, m in ((300, 300), (512, 256), (200, 600)):
        x, y = torch.rand(2, n, 3, generator=g), torch.rand(2, m, 3, generator=g)
        ref, _ = OO.emd_approxmatch_cost(x, y)
        np.testing.assert_allclose(ops.emd_approx(x.cuda(), y.cuda()).cpu().numpy(), ref.numpy(), rtol=2e-3)
    nograd = ops.earth_mover_distance_nograd(a.cuda().transpose(1, 2), b.cuda().transpose(1, 2))
    np.testing.assert_allclose(nograd.cpu().numpy(), ref_ops["emd_cost"] / a.shape[1], rtol=2e-3)


@pytest.mark.parametrize("N,K,Q", [(5000, 2048, 7), (777, 777, 3), (20000, 300, 5), (4096, 1, 4)])
def test_knn_points_matches_oracle_bitwise(N, K, Q):
    """Radix-select + sort kNN == oracle (ascending squared distance, ties by lower index), indices and distances."""
    from oracle import ops as OO
    from p2pb_b200 import ops

    g = torch.Generator().manual_seed(N + K)
    pts = torch.randn(N, 3, generator=g)
    if N == 777:                       # lattice: almost every distance ties
        pts = torch.randint(0, 5, (N, 3), generator=g).float() * 0.5
    q = pts[torch.randperm(N, generator=g)[:Q]].contiguous()
    idx, dist = ops.knn_points(q.cuda(), pts.cuda(), K, return_dist=True)
    ridx, rdist = OO.knn_points(q, pts, K)
    assert torch.equal(idx.cpu(), ridx)
    assert torch.equal(dist.cpu(), rdist)


@pytest.mark.parametrize("B,N,M,snap", [(3, 8192, 2048, None), (2, 4000, 700, 16), (1, 30000, 900, None), (1, 150000, 600, 64),
                                        (16, 8192, 256, None)])
def test_fps_cluster_kernel_bit_exact(B, N, M, snap):
    """Thread-block-cluster FPS (points split over 8 / 16 CTAs, DSMEM candidate exchange): indices identical to the
    oracle (reference tie-break, heavy ties on the snapped lattices) and to the one-CTA-per-cloud kernels; fused centres."""
    from oracle import ops as OO
    from p2pb_b200 import ops
    from p2pb_b200._lib import lib

    coords = cloud(B, N, seed=N + M, snap=snap)
    cd = coords.cuda()
    lib().p2pb_fps_set_cluster(2)          # clusters for the patch sizes too (default: whole clouds only)
    idx, centers = ops.furthest_point_sampling(cd, M, return_centers=True)
    assert torch.equal(idx.cpu(), OO.furthest_point_sampling_forward(coords, M))
    assert torch.equal(centers, torch.gather(cd, 2, idx.long()[:, None, :].expand(-1, 3, -1)))
    lib().p2pb_fps_set_cluster(0)
    try:
        assert torch.equal(ops.furthest_point_sampling(cd, M), idx)
    finally:
        lib().p2pb_fps_set_cluster(1)


@pytest.mark.parametrize("B,N,M,snap", [(1, 400000, 300, None), (1, 1000003, 64, 32), (2, 250000, 100, None)])
def test_fps_grid_kernel_bit_exact(B, N, M, snap):
    """Whole-GPU cooperative FPS (one CTA per SM, grid barrier per iteration) for room-sized clouds: indices identical to the oracle
    (reference tie-break; heavy ties on the snapped lattice) and to the one-CTA global-memory kernel."""
    from oracle import ops as OO
    from p2pb_b200 import ops
    from p2pb_b200._lib import lib

    coords = cloud(B, N, seed=N + M, snap=snap)
    cd = coords.cuda()
    idx = ops.furthest_point_sampling(cd, M)
    assert torch.equal(idx.cpu(), OO.furthest_point_sampling_forward(coords, M))
    lib().p2pb_fps_set_cluster(0)          # one CTA per cloud, distances in global memory
    try:
        assert torch.equal(ops.furthest_point_sampling(cd, M), idx)
    finally:
        lib().p2pb_fps_set_cluster(1)


@pytest.mark.parametrize("N,P,radius", [(50000, 37, 0.2), (3000, 5, 10.0), (4097, 9, 0.0)])
def test_radius_query_matches_oracle(N, P, radius):
    """Device radius query (count + index-ordered fill) == oracle CSR, bit for bit (incl. 'everything' and 'only the centre')."""
    from oracle import ops as OO
    from p2pb_b200 import ops

    g = torch.Generator().manual_seed(N)
    pts = torch.rand(N, 3, generator=g)
    ctr = pts[torch.randperm(N, generator=g)[:P]].contiguous()
    off, idx = ops.radius_query(ctr.cuda(), pts.cuda(), radius)
    roff, ridx = OO.radius_query(ctr, pts, radius)
    assert torch.equal(off.cpu(), roff) and torch.equal(idx.cpu(), ridx)
