"""GPU tests of the tcgen05 GEMM / implicit-GEMM conv kernels against fp64 torch references.
TF32 tolerance (10-bit mantissa operands, fp32 accumulate): max |err| <= 2e-3 * max |ref| -- the arithmetic class of
the reference's own cuDNN-TF32 convolutions."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-3


def _close(out, ref):
    err = (out.double() - ref).abs().max().item()
    assert err <= TOL * ref.abs().max().item() + 1e-6, (err, ref.abs().max().item())


@pytest.mark.parametrize("M,K,N", [(128, 32, 16), (256, 64, 64), (1000, 96, 32), (64, 1024, 256), (4096, 128, 512),
                                   (300, 32, 48), (131072, 64, 128), (8, 512, 384)])
def test_gemm_rows_vs_fp64(M, K, N):
    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    out = dense.gemm_rows([A], W, bias)
    _close(out, A.double() @ W.double().t() + bias.double())


def test_gemm_rows_segments_bias2_stats():
    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(0)
    B, rows = 6, 256
    M = B * rows
    wide = torch.randn(M, 160, device="cuda", generator=g)      # segment 0 = columns 32..95 of a wider buffer
    A0 = wide[:, 32:96]
    A1 = torch.randn(M, 32, device="cuda", generator=g)
    A2 = torch.randn(M, 96, device="cuda", generator=g)
    N = 64
    W = torch.randn(N, 64 + 32 + 96, device="cuda", generator=g) / 14.0
    bias = torch.randn(N, device="cuda", generator=g)
    bias2 = torch.randn(B, N, device="cuda", generator=g)
    stats = torch.zeros(dense.num_m_tiles(M), N, 2, device="cuda")
    out = torch.full((M, 80), 7.0, device="cuda")            # ldd > N: columns beyond N stay untouched
    dense.gemm_rows([A0, A1, A2], W, bias, bias2, rows, out=out[:, :N], stats=stats)
    ref = torch.cat([A0, A1, A2], 1).double() @ W.double().t() + bias.double() + bias2.double().repeat_interleave(rows, 0)
    _close(out[:, :N], ref)
    assert torch.all(out[:, N:] == 7.0)
    s = stats.double().view(B, rows // 128, N, 2).sum(1)
    o = out[:, :N].double().view(B, rows, N)
    assert torch.allclose(s[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("B,r,cin,cout", [(2, 8, 32, 16), (1, 16, 64, 64), (2, 32, 32, 32), (3, 8, 256, 256), (1, 16, 128, 128),
                                          (1, 32, 64, 64), (2, 8, 192, 128)])
def test_conv3d_vs_fp64(B, r, cin, cout):
    import torch.nn.functional as F

    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(r + cin)
    x = torch.randn(B, cin, r, r, r, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    grid = x.permute(0, 2, 3, 4, 1).contiguous()               # channels-last [B, r, r, r, cin]
    stats = torch.zeros(B * r ** 3 // 128, cout, 2, device="cuda")
    out = dense.conv3d_cl(grid, dense.pack_conv3d_weight(w, cin), bias, B, r, cin, cout, stats=stats)
    ref = F.conv3d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1).reshape(B * r ** 3, cout)
    _close(out, ref)
    s = stats.double().view(B, -1, cout, 2).sum(1)
    assert torch.allclose(s[..., 0], out.double().view(B, -1, cout).sum(1), rtol=1e-4, atol=1e-2)


def test_conv3d_padded_channels_and_permutation():
    """Cin=35 padded to 64 with a channel permutation (features first, xyz last), as the engine stores SA0's input."""
    import torch.nn.functional as F

    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(3)
    B, r, cin, cout, cp = 1, 8, 35, 32, 64
    x = torch.randn(B, cin, r, r, r, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / 30.0
    perm = list(range(3, 35)) + [0, 1, 2]
    grid = torch.zeros(B, r, r, r, cp, device="cuda")
    grid[..., :cin] = x[:, perm].permute(0, 2, 3, 4, 1)
    out = dense.conv3d_cl(grid, dense.pack_conv3d_weight(w, cp, perm), None, B, r, cp, cout)
    ref = F.conv3d(x.double(), w.double(), None, padding=1).permute(0, 2, 3, 4, 1).reshape(-1, cout)
    _close(out, ref)


@pytest.mark.parametrize("B,r,cin,cout", [(1, 8, 32, 32), (2, 16, 64, 64), (1, 32, 32, 32), (2, 32, 64, 64), (1, 16, 128, 128),
                                          (3, 16, 128, 64), (1, 32, 64, 32)])
def test_conv3d_halo_vs_fp64(B, r, cin, cout):
    """Halo-reuse conv (chunk-planar padded input, un-swizzled A descriptors, stationary weight slabs)."""
    import torch.nn.functional as F

    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(r + cin + cout)
    x = torch.randn(B, cin, r, r, r, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    X = dense.dense_to_padded(x.permute(0, 2, 3, 4, 1).contiguous(), r)
    _, _, tps = dense.halo_layout(r)
    stats = torch.zeros(B * tps, cout, 2, device="cuda")
    out = torch.full((B * r ** 3, cout), float("nan"), device="cuda")
    dense.conv3d_halo(X, dense.pack_conv3d_weight(w, cin), bias, B, r, cin, cout, out=out, stats=stats)
    ref = F.conv3d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1).reshape(B * r ** 3, cout)
    _close(out, ref)
    s = stats.double().view(B, tps, cout, 2).sum(1)
    o = out.double().view(B, -1, cout)
    assert torch.allclose(s[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)
