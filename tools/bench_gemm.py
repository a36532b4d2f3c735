"""Dev tool: the rows-GEMMs / r = 8 convs of one PVDS evaluation (64 patches) on the persistent tcgen05 kernel -- fp32-stored
(kind::tf32) and IEEE-half operands (kind::f16), CTA pairs vs independent CTAs -- next to cuBLAS (torch.nn.functional.linear with
TF32 and with fp16 inputs).  Prints a markdown table (profiles/r02_gemm_bench.md)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
from p2pb_b200._lib import lib


def timeit(fn, n=20):
    """fn(i) launches on buffer set i; the caller rotates enough sets that no launch finds its operands or the lines it writes in
    the 126 MB L2 (a 67 MB output re-written in place never reaches HBM and flatters an HBM-bound shape by up to 2x)."""
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(3 + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3       # us


def n_sets(bytes_per_set):
    return max(2, -(-400_000_000 // int(bytes_per_set)))     # > 3 x L2 between two uses of a set


HBM, TENSOR16 = 6550.7e9, 1602.9e12
print("| M | K | N | tf32 pairs | tf32 single | half pairs | half stats-only | cuBLAS tf32 | cuBLAS fp16 | half: TFLOP/s | half: GB/s (frac of 6551) | bound |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for (M, K, N) in [(131072, 64, 128), (131072, 128, 256), (131072, 256, 512), (131072, 512, 1024), (1048576, 64, 32),
                  (1048576, 64, 64), (262144, 128, 64), (262144, 64, 128), (65536, 192, 128), (65536, 128, 256), (131072, 256, 128),
                  (131072, 128, 128), (131072, 64, 64), (8192, 384, 256), (2048, 832, 512)]:
    R = n_sets(4.0 * M * K + 4.0 * M * N)
    W = torch.randn(N, K, device="cuda") / K ** 0.5; Wh = W.half()
    As = [torch.randn(M, K, device="cuda") for _ in range(R)]; Ahs = [a.half() for a in As]
    outs = [torch.empty(M, N, device="cuda") for _ in range(R)]; outh = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(R)]
    bias = torch.randn(N, device="cuda")
    stats = torch.zeros(dense.num_stat_blocks(M), N, 2, device="cuda")
    colmm = torch.zeros(dense.num_stat_blocks(M), N, 2, device="cuda")
    t_pair = timeit(lambda i: dense.gemm_rows([As[i % R]], W, bias, out=outs[i % R], stats=stats))
    lib().p2pb_gemm_tune(32)
    t_single = timeit(lambda i: dense.gemm_rows([As[i % R]], W, bias, out=outs[i % R], stats=stats))
    lib().p2pb_gemm_tune(0)
    t_half = timeit(lambda i: dense.gemm_rows([Ahs[i % R]], Wh, bias, out=outs[i % R], stats=stats))
    t_half_so = timeit(lambda i: dense.gemm_rows([Ahs[i % R]], Wh, bias, stats=stats, colmm=colmm, store=False))
    torch.backends.cuda.matmul.allow_tf32 = True
    t_cb = timeit(lambda i: torch.addmm(bias, As[i % R], W.t(), out=outs[i % R]))
    bh = bias.half()
    t_cbh = timeit(lambda i: torch.addmm(bh, Ahs[i % R], Wh.t(), out=outh[i % R]))
    del As, Ahs, outs, outh
    fl = 2.0 * M * K * N
    by_half = 2.0 * M * K + 4.0 * M * N
    tf, gb = fl / t_half / 1e6, by_half / t_half / 1e3
    bound = "tensor" if fl / TENSOR16 > by_half / HBM else "hbm"
    print(f"| {M} | {K} | {N} | {t_pair:.1f} | {t_single:.1f} | {t_half:.1f} | {t_half_so:.1f} | {t_cb:.1f} | {t_cbh:.1f} | {tf:.0f} | {gb:.0f} ({gb / 6550.7:.2f}) | {bound} |")

print()
print("| conv r=8 (64 patches) | tf32 | half | half TFLOP/s |")
print("|---|---|---|---|")
for (B, r, cin, cout) in [(64, 8, 256, 256), (64, 8, 256, 128), (64, 8, 128, 128)]:
    grid = torch.randn(B, r, r, r, cin, device="cuda")
    w = torch.randn(cout, 27 * cin, device="cuda") / (27 * cin) ** 0.5
    bias = torch.randn(cout, device="cuda"); out = torch.empty(B * r ** 3, cout, device="cuda")
    stats = torch.zeros(B * r ** 3 // 32, cout, 2, device="cuda")
    t32 = timeit(lambda i: dense.conv3d_cl(grid, w, bias, B, r, cin, cout, out=out, stats=stats))
    gh, wh = grid.half(), w.half()
    t16 = timeit(lambda i: dense.conv3d_cl(gh, wh, bias, B, r, cin, cout, out=out, stats=stats))
    fl = 2.0 * B * r ** 3 * 27 * cin * cout
    print(f"| {cin}->{cout} | {t32:.1f} | {t16:.1f} | {fl / t16 / 1e6:.0f} |")
