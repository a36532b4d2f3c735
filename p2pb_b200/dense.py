"""Torch-tensor front end of the tcgen05 GEMM / implicit-GEMM conv kernels (``csrc/gemm_persist.cu``, ``csrc/conv_halo.cu``).

Channels-last fp32 activations ("rows": ``[M, C]`` with C padded to a multiple of 32); weights pre-packed K-major
``[N, K]``.  Kernels run on torch's current stream; CUDA only."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from ._lib import P2PBError, call

_vp = ctypes.c_void_p


def _p(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _s():
    return _vp(torch.cuda.current_stream().cuda_stream)


def num_m_tiles(M: int) -> int:
    return (M + 127) // 128


def num_stat_blocks(M: int) -> int:
    """Rows of the ``stats`` / ``colmm`` outputs of gemm_rows / conv3d_cl: one per 32-row block of every 128-row tile."""
    return 4 * num_m_tiles(M)


def gemm_rows(segs: Sequence[torch.Tensor], W: torch.Tensor, bias: Optional[torch.Tensor] = None,
              bias2: Optional[torch.Tensor] = None, rows_per_sample: int = 0, out: Optional[torch.Tensor] = None,
              stats: Optional[torch.Tensor] = None, ks: Optional[Sequence[int]] = None,
              colmm: Optional[torch.Tensor] = None, store: bool = True) -> Optional[torch.Tensor]:
    """D[M,N] = cat(segs, dim=1)[M,K] @ W[N,K]^T + bias + bias2[m // rows_per_sample].

    ``segs``: 1-3 fp32 2-D tensors with unit inner stride (column slices of wider buffers are fine); ``ks`` optionally
    restricts the number of leading columns used from each segment.  ``stats`` / ``colmm`` [num_stat_blocks(M), N, 2]
    receive the column (sum, sum^2) / (max, min) of every 32-row block; ``store=False`` skips D altogether (statistics-only GEMM)."""
    assert 1 <= len(segs) <= 3
    M = segs[0].shape[0]
    N = W.shape[0]
    f16 = W.dtype == torch.float16          # IEEE-half operands -> p2pb_gemm_rows_f16 (fp32 bias / accumulate / output)
    a = []
    for i in range(3):
        if i < len(segs):
            t = segs[i]
            if not (t.is_cuda and t.dtype == W.dtype and t.dim() == 2 and t.stride(1) == 1 and t.shape[0] == M):
                raise P2PBError(f"gemm_rows: segment {i} must be a CUDA {W.dtype} [M,K] tensor with unit inner stride")
            k = t.shape[1] if ks is None else ks[i]
            a += [_p(t), int(k), int(t.stride(0))]
        else:
            a += [_vp(0), 0, 0]
    if out is None and store:
        out = torch.empty((M, N), dtype=torch.float32, device=W.device)
    if not store:
        out = None
    assert (out is None or out.stride(1) == 1) and W.is_contiguous()
    with torch.cuda.device(W.device):
        call("p2pb_gemm_rows_f16" if f16 else "p2pb_gemm_rows_ex", *a, _p(W), _p(bias), _p(bias2), int(rows_per_sample), _p(out),
             int(out.stride(0)) if out is not None else 0, _p(stats), _p(colmm), int(M), int(N), _s())
    return out


def conv3d_cl(grid: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], B: int, r: int, cin: int, cout: int,
              out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3x3 / stride 1 / pad 1 conv on a channels-last grid [B, r, r, r, cin] -> rows [B*r^3, cout]."""
    if out is None:
        out = torch.empty((B * r ** 3, cout), dtype=torch.float32, device=grid.device)
    assert grid.is_contiguous() and W.is_contiguous() and out.stride(1) == 1 and grid.dtype == W.dtype
    with torch.cuda.device(grid.device):
        call("p2pb_conv3d_cl_f16" if W.dtype == torch.float16 else "p2pb_conv3d_cl", _p(grid), _p(W), _p(bias), _p(out), int(out.stride(0)), _p(stats), int(B), int(r), int(cin),
             int(cout), _s())
    return out


def pack_conv3d_weight(w: torch.Tensor, cin_pad: int, perm: Optional[Sequence[int]] = None) -> torch.Tensor:
    """[Cout, Cin, 3, 3, 3] (reference layout) -> [Cout, 27*cin_pad], k = ((kx*3+ky)*3+kz)*cin_pad + c.
    ``perm[c_new] = c_old`` optionally reorders input channels (the engine stores features before coordinates)."""
    cout, cin = w.shape[:2]
    if perm is not None:
        w = w[:, list(perm)]
    p = torch.zeros((cout, 27, cin_pad), dtype=torch.float32, device=w.device)
    p[:, :, :cin] = w.reshape(cout, cin, 27).permute(0, 2, 1)
    return p.reshape(cout, 27 * cin_pad).contiguous()


# ---- halo-reuse conv (csrc/conv_halo.cu): zero-bordered row-major input -------------------------------------------
def halo_layout(r: int, cout: Optional[int] = None, f16: bool = True, cin: int = 64):
    """(rows per sample P^3, slack rows after the last sample, tiles per sample) of the padded layout.  The tile count (= rows
    of the ``stats`` output per sample) depends on the kernel variant: pass ``cin`` (padded) / ``cout`` / ``f16`` of the convolution."""
    from ._lib import lib

    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    call("p2pb_conv_halo_layout", int(r), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    tiles = c.value if cout is None else int(lib().p2pb_conv_halo_tiles(int(r), int(cin), int(cout), int(bool(f16))))
    return a.value, b.value, tiles


def alloc_padded(B: int, C: int, r: int, device, dtype=torch.float32) -> torch.Tensor:
    """Zero-initialised X[B*(r+2)^3 + slack, C]; border rows stay zero forever (kernels only write interior rows)."""
    P3, slack, _ = halo_layout(r)
    return torch.zeros(B * P3 + slack, C, dtype=dtype, device=device)


def dense_to_padded(grid: torch.Tensor, r: int) -> torch.Tensor:
    """Test helper (torch ops): channels-last dense grid [B, r, r, r, C] -> padded row-major layout (dtype kept)."""
    B, C = grid.shape[0], grid.shape[-1]
    P = r + 1
    X = alloc_padded(B, C, r, grid.device, grid.dtype)
    X[: B * P ** 3].view(B, P, P, P, C)[:, 1:, 1:, 1:, :] = grid
    return X


def conv3d_halo(X: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], B: int, r: int, cin: int, cout: int,
                out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
                cin_valid: Optional[int] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty((B * r ** 3, cout), dtype=torch.float32, device=X.device)
    assert X.is_contiguous() and X.shape[1] == cin and X.dtype == W.dtype
    entry = "p2pb_conv3d_halo_f16" if X.dtype == torch.float16 else "p2pb_conv3d_halo_ex"
    with torch.cuda.device(X.device):
        call(entry, _p(X), _p(W), _p(bias), _p(out), int(out.stride(0)), _p(stats), int(B), int(r), int(cin),
             int(cin if cin_valid is None else cin_valid), int(cout), _s())
    return out
