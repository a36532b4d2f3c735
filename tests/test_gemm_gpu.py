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


@pytest.mark.parametrize("M,K,N", [(128, 32, 32), (256, 64, 64), (1000, 96, 32), (64, 1024, 256), (4096, 128, 512),
                                   (300, 32, 96), (131072, 64, 128), (8, 512, 384)])
def test_gemm_rows_vs_fp64(M, K, N):
    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    out = dense.gemm_rows([A], W, bias)
    _close(out, A.double() @ W.double().t() + bias.double())


@pytest.mark.parametrize("mode", [0, 4, 32])
@pytest.mark.parametrize("M,K,N", [(128 * 301 - 50, 96, 256), (128 * 300, 512, 1024), (40000, 64, 128), (5000, 160, 96)])
def test_gemm_persistent_cluster_minmax(M, K, N, mode):
    """Persistent kernel with cta_group::2 CTA pairs, M = 256, for the big shapes (mode 0; odd tile count = unpaired tail), with
    2-CTA multicast clusters (mode 4) and with independent CTAs only (mode 32), p2pb_gemm_tune: results, GroupNorm partials,
    column (max, min) and the statistics-only variant (no D)."""
    from p2pb_b200 import dense
    from p2pb_b200._lib import lib

    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.double() @ W.double().t() + bias.double()
    T = dense.num_stat_blocks(M)
    assert lib().p2pb_gemm_tune(mode) == 0
    try:
        stats = torch.zeros(T, N, 2, device="cuda")
        colmm = torch.zeros(T, N, 2, device="cuda")
        out = dense.gemm_rows([A], W, bias, stats=stats, colmm=colmm)
        _close(out, ref)
        assert torch.allclose(stats[..., 0].double().sum(0), out.double().sum(0), rtol=1e-4, atol=1e-2)
        assert torch.allclose(stats[..., 1].double().sum(0), (out.double() ** 2).sum(0), rtol=1e-4, atol=1e-2)
        if colmm is not None:
            assert torch.equal(colmm[..., 0].max(0).values, out.max(0).values)
            assert torch.equal(colmm[..., 1].min(0).values, out.min(0).values)
            stats2 = torch.zeros_like(stats)
            colmm2 = torch.zeros_like(colmm)
            assert dense.gemm_rows([A], W, bias, stats=stats2, colmm=colmm2, store=False) is None
            assert torch.equal(stats2, stats) and torch.equal(colmm2, colmm)
    finally:
        lib().p2pb_gemm_tune(0)


def test_gmax_minmax_matches_dense_pass():
    """max over rows of Swish(x*A+B) from the column (max, min) == the dense affine_act + gmax pass."""
    import ctypes

    from p2pb_b200 import dense
    from p2pb_b200._lib import call

    g = torch.Generator(device="cuda").manual_seed(5)
    B, rows, K, N = 3, 512, 64, 128
    A = torch.randn(B * rows, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / 8
    colmm = torch.zeros(B * rows // 32, N, 2, device="cuda")
    out = dense.gemm_rows([A], W, None, colmm=colmm)
    ca = torch.randn(B, N, device="cuda", generator=g)
    cb = torch.randn(B, N, device="cuda", generator=g)
    got = torch.empty(B, N, device="cuda")
    vp = ctypes.c_void_p
    call("p2pb_gmax_minmax", vp(colmm.data_ptr()), rows // 32, B, N, vp(ca.data_ptr()), vp(cb.data_ptr()), 1, vp(got.data_ptr()),
         vp(torch.cuda.current_stream().cuda_stream))
    y = out.view(B, rows, N) * ca[:, None] + cb[:, None]
    ref = (y * torch.sigmoid(y)).max(1).values
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)


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
    stats = torch.zeros(dense.num_stat_blocks(M), N, 2, device="cuda")
    out = torch.full((M, 80), 7.0, device="cuda")            # ldd > N: columns beyond N stay untouched
    dense.gemm_rows([A0, A1, A2], W, bias, bias2, rows, out=out[:, :N], stats=stats)
    ref = torch.cat([A0, A1, A2], 1).double() @ W.double().t() + bias.double() + bias2.double().repeat_interleave(rows, 0)
    _close(out[:, :N], ref)
    assert torch.all(out[:, N:] == 7.0)
    s = stats.double().view(B, rows // 32, N, 2).sum(1)
    o = out[:, :N].double().view(B, rows, N)
    assert torch.allclose(s[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("B,r,cin,cout", [(2, 8, 32, 32), (1, 16, 64, 64), (2, 32, 32, 32), (3, 8, 256, 256), (1, 16, 128, 128),
                                          (1, 32, 64, 64), (2, 8, 192, 128)])
def test_conv3d_vs_fp64(B, r, cin, cout):
    import torch.nn.functional as F

    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(r + cin)
    x = torch.randn(B, cin, r, r, r, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    grid = x.permute(0, 2, 3, 4, 1).contiguous()               # channels-last [B, r, r, r, cin]
    stats = torch.zeros(B * r ** 3 // 32, cout, 2, device="cuda")
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
                                          (3, 16, 128, 64), (1, 32, 64, 32), (8, 16, 64, 64), (9, 16, 128, 128), (3, 32, 64, 32),
                                          (15, 16, 32, 32)])
def test_conv3d_halo_vs_fp64(B, r, cin, cout):
    """Halo-reuse conv (row-shifted windows, sub-slab weight ring).  The larger batches (>= 296 tiles) run as
    cta_group::2 CTA pairs (M = 256), including odd tile counts (dummy slot) -- (9,16,..): 369 tiles, (15,16,..): 615."""
    import torch.nn.functional as F

    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(r + cin + cout)
    x = torch.randn(B, cin, r, r, r, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    X = dense.dense_to_padded(x.permute(0, 2, 3, 4, 1).contiguous(), r)
    _, _, tps = dense.halo_layout(r, cout, f16=False, cin=cin)
    stats = torch.zeros(B * tps, cout, 2, device="cuda")
    out = torch.full((B * r ** 3, cout), float("nan"), device="cuda")
    dense.conv3d_halo(X, dense.pack_conv3d_weight(w, cin), bias, B, r, cin, cout, out=out, stats=stats)
    ref = F.conv3d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1).reshape(B * r ** 3, cout)
    _close(out, ref)
    s = stats.double().view(B, tps, cout, 2).sum(1)
    o = out.double().view(B, -1, cout)
    assert torch.allclose(s[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("B,r,cin,cin_valid,cout", [(1, 16, 64, 64, 64), (2, 32, 64, 35, 32), (9, 16, 128, 128, 128), (3, 32, 64, 32, 32),
                                                    (15, 16, 128, 100, 64), (5, 32, 64, 64, 64), (9, 16, 192, 192, 64), (12, 32, 64, 35, 32),
                                                    (1, 16, 64, 64, 32), (64, 16, 64, 64, 64)])
def test_conv3d_halo_half_operands_vs_fp64(B, r, cin, cin_valid, cout):
    """IEEE-half operand variant (kind::f16, 64 channels per chunk, K-valid MMA skipping), un-paired and CTA-pair
    launches: compared with the fp64 convolution of the SAME half-rounded operands at the tf32 tolerance (half and
    tf32 share the 10-bit mantissa) and with the fp32-operand kernel."""
    import torch.nn.functional as F

    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(r + cin + cout)
    x = torch.randn(B, cin, r, r, r, device="cuda", generator=g)
    x[:, cin_valid:] = 0
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5
    w[:, cin_valid:] = 0
    bias = torch.randn(cout, device="cuda", generator=g)
    xh, wh = x.half(), w.half()
    X = dense.dense_to_padded(xh.permute(0, 2, 3, 4, 1).contiguous(), r)
    _, _, tps = dense.halo_layout(r, cout, f16=True, cin=cin)       # Cout = 32 / (Cout = 64, Cin >= 192): the dz-stacked form
    stats = torch.zeros(B * tps, cout, 2, device="cuda")
    out = torch.full((B * r ** 3, cout), float("nan"), device="cuda")
    dense.conv3d_halo(X, dense.pack_conv3d_weight(wh.float(), cin).half(), bias, B, r, cin, cout, out=out, stats=stats,
                      cin_valid=cin_valid)
    ref = F.conv3d(xh.double(), wh.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1).reshape(B * r ** 3, cout)
    err = (out.double() - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item() + 1e-6, err          # exact products, fp32 accumulation
    ref32 = F.conv3d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1).reshape(B * r ** 3, cout)
    _close(out, ref32)                                               # vs the un-rounded operands: tf32-class error
    s = stats.double().view(B, tps, cout, 2).sum(1)
    o = out.double().view(B, -1, cout)
    assert torch.allclose(s[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("M,K,N", [(128 * 301 - 50, 128, 256), (40000, 64, 128), (5000, 192, 96), (131072, 512, 1024)])
def test_gemm_rows_half_operands(M, K, N):
    """IEEE-half operand GEMM (kind::f16; single CTAs and cta_group::2 pairs): exact products + fp32 accumulation vs the fp64
    product of the same half-rounded operands; GroupNorm partials, column max/min, statistics-only variant."""
    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    T = dense.num_stat_blocks(M)
    stats = torch.zeros(T, N, 2, device="cuda")
    colmm = torch.zeros(T, N, 2, device="cuda")
    out = dense.gemm_rows([A], W, bias, stats=stats, colmm=colmm)
    ref = A.double() @ W.double().t() + bias.double()
    err = (out.double() - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item() + 1e-6, err
    assert torch.allclose(stats[..., 0].double().sum(0), out.double().sum(0), rtol=1e-4, atol=1e-2)
    assert torch.equal(colmm[..., 0].max(0).values, out.max(0).values)
    assert torch.equal(colmm[..., 1].min(0).values, out.min(0).values)
    stats2, colmm2 = torch.zeros_like(stats), torch.zeros_like(colmm)
    assert dense.gemm_rows([A], W, bias, stats=stats2, colmm=colmm2, store=False) is None
    assert torch.equal(stats2, stats) and torch.equal(colmm2, colmm)


@pytest.mark.parametrize("B,r,cin,cout", [(3, 8, 256, 256), (64, 8, 128, 128), (2, 16, 64, 64)])
def test_conv3d_cl_half_operands(B, r, cin, cout):
    """Per-tap implicit-GEMM conv with IEEE-half grid / weights vs fp64 on the same rounded operands."""
    import torch.nn.functional as F

    from p2pb_b200 import dense

    g = torch.Generator(device="cuda").manual_seed(r + cin)
    x = torch.randn(B, cin, r, r, r, device="cuda", generator=g).half()
    w = (torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5).half()
    bias = torch.randn(cout, device="cuda", generator=g)
    grid = x.permute(0, 2, 3, 4, 1).contiguous()
    stats = torch.zeros(B * r ** 3 // 32, cout, 2, device="cuda")
    out = dense.conv3d_cl(grid, dense.pack_conv3d_weight(w.float(), cin).half(), bias, B, r, cin, cout, stats=stats)
    ref = F.conv3d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1).reshape(B * r ** 3, cout)
    err = (out.double() - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item() + 1e-6, err
    s = stats.double().view(B, -1, cout, 2).sum(1)
    assert torch.allclose(s[..., 0], out.double().view(B, -1, cout).sum(1), rtol=1e-4, atol=1e-2)


def test_gemm_rejects_shapes_outside_the_envelope():
    """N must be a multiple of 32 (one TMEM column block); there is no second kernel to fall back to -- the call raises."""
    from p2pb_b200 import dense
    from p2pb_b200._lib import P2PBError

    A = torch.randn(128, 32, device="cuda")
    with pytest.raises(P2PBError, match="multiple of 32"):
        dense.gemm_rows([A], torch.randn(16, 32, device="cuda"), None)


def test_step_vectors_se_excite_head_bridge_match_torch():
    """The three small fused kernels of one sampling step against plain torch fp32."""
    import ctypes

    from p2pb_b200._lib import call

    vp = ctypes.c_void_p
    p = lambda t: vp(t.data_ptr()) if t is not None else vp(0)
    s = vp(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(3)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    B, E = 5, 64
    sin, w0, b0, w2, b2 = rn(B, E), rn(E, E) / 8, rn(E), rn(E, E) / 8, rn(E)
    folds = [rn(c, E) / 8 for c in (64, 128, 37, 256)]
    bufs = [torch.zeros(B, w.shape[0], device="cuda") for w in folds]
    ptr = torch.tensor([b.data_ptr() + 4 * o for b, w in zip(bufs, folds) for o in range(w.shape[0])], dtype=torch.int64, device="cuda")
    stride = torch.tensor([w.shape[0] for w in folds for _ in range(w.shape[0])], dtype=torch.int32, device="cuda")
    wall = torch.cat(folds, 0).contiguous()
    temb = torch.zeros(B, E, device="cuda")
    call("p2pb_step_vectors", p(sin), E, p(w0), p(b0), p(w2), p(b2), B, E, p(wall), wall.shape[0], p(ptr), p(stride), p(temb), s)
    t_ref = torch.nn.functional.linear(torch.nn.functional.leaky_relu(torch.nn.functional.linear(sin, w0, b0), 0.1), w2, b2)
    assert torch.allclose(temb, t_ref, atol=1e-5)
    for b, w in zip(bufs, folds):
        assert torch.allclose(b, t_ref @ w.t(), atol=1e-4)
    # shared noise level (row stride 0), no folds
    temb0 = torch.zeros(B, E, device="cuda")
    call("p2pb_step_vectors", p(sin[:1]), 0, p(w0), p(b0), p(w2), p(b2), B, E, vp(0), 0, vp(0), vp(0), p(temb0), s)
    assert torch.allclose(temb0, t_ref[:1].expand(B, E), atol=1e-5)
    # SE gate
    C, Hd = 256, 32
    ym, v0, v2 = rn(B, C), rn(Hd, C) / 16, rn(C, Hd) / 6
    se = torch.zeros(B, C, device="cuda")
    call("p2pb_se_excite", p(ym), p(v0), p(v2), B, C, Hd, p(se), s)
    assert torch.allclose(se, torch.sigmoid(torch.relu(ym @ v0.t()) @ v2.t()), atol=1e-5)
    # classifier tail + bridge update
    N, C = 777, 128
    raw, A, Bc, W, bias = rn(B * N, C), rn(B, C), rn(B, C), rn(3, C) / 11, rn(3)
    xt, coef = rn(B, 3, N), torch.tensor([0.7, 0.3, 0.65], device="cuda")
    h = torch.nn.functional.silu(raw.view(B, N, C) * A[:, None] + Bc[:, None])
    eps_ref = (h @ W.t() + bias).permute(0, 2, 1)                                   # [B,3,N]
    eps = torch.zeros(B * N, 16, device="cuda")
    call("p2pb_head_bridge", p(raw), C, p(A), p(Bc), p(W), p(bias), B, C, N, vp(0), vp(0), 0, vp(0), vp(0), p(eps), 16, s)
    got = eps[:, :3].view(B, N, 3).permute(0, 2, 1)
    assert torch.allclose(got, eps_ref, atol=2e-5, rtol=1e-5)
    for clip in (0, 1):
        xn, x0 = torch.zeros_like(xt), torch.zeros_like(xt)
        call("p2pb_head_bridge", p(raw), C, p(A), p(Bc), p(W), p(bias), B, C, N, p(xt * 4), p(coef), clip, p(xn), p(x0), vp(0), 0, s)
        p0 = xt * 4 - coef[0] * got
        if clip:
            p0 = p0.clamp(-3.0, 3.0)
        assert torch.equal(x0, p0) and torch.equal(xn, coef[1] * p0 + coef[2] * (xt * 4))   # same fp32 op order as p2pb.py:155-165,190-213


@pytest.mark.parametrize("B,r,cin,cout", [(3, 32, 64, 64), (5, 16, 64, 64), (2, 16, 128, 64)])
def test_conv3d_halo_stacked_form_forced_equals_default(B, r, cin, cout):
    """p2pb_conv_halo_tune(0, 0, 200) forces the dz-stacked form (N = 3 Cout per MMA, +-1 lane shift in the epilogue) on the shapes
    the default dispatch keeps un-stacked: same convolution, fp32-accumulation-order differences only."""
    from p2pb_b200 import dense
    from p2pb_b200._lib import lib

    g = torch.Generator(device="cuda").manual_seed(r + cin + cout)
    x = torch.randn(B, r, r, r, cin, device="cuda", generator=g).half()
    w = (torch.randn(cout, 27 * cin, device="cuda", generator=g) / (27 * cin) ** 0.5).half()
    bias = torch.randn(cout, device="cuda", generator=g)
    X = dense.dense_to_padded(x, r)
    outs = []
    for G in (0, 200, 100):
        assert lib().p2pb_conv_halo_tune(0, 0, G) == 0
        try:
            _, _, tps = dense.halo_layout(r, cout, True, cin=cin)
            st = torch.zeros(B * tps, cout, 2, device="cuda")
            out = torch.full((B * r ** 3, cout), float("nan"), device="cuda")
            dense.conv3d_halo(X, w, bias, B, r, cin, cout, out=out, stats=st)
            outs.append((out, st.view(B, tps, cout, 2).sum(1)))
        finally:
            lib().p2pb_conv_halo_tune(0, 0, 0)
    for out, st in outs[1:]:
        assert torch.allclose(out, outs[0][0], rtol=1e-4, atol=1e-4)
        assert torch.allclose(st, outs[0][1], rtol=1e-4, atol=1e-2)
