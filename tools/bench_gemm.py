"""Dev tool: time the rows-GEMMs / r=8 convs of the PVDS evaluation (B=64): legacy one-tile-per-CTA kernel (mode 2),
persistent kernel without clusters (mode 4) and persistent kernel with 2-CTA multicast clusters (mode 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
from p2pb_b200._lib import lib


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("rows GEMM: us for legacy | persist | persist+cluster | persist stats-only(minmax) | torch fp32-linear(tf32 off)")
for (M, K, N) in [(131072, 32, 128), (131072, 128, 256), (131072, 256, 512), (131072, 512, 1024), (1048576, 64, 32),
                  (1048576, 32, 64), (262144, 96, 64), (262144, 64, 128), (65536, 160, 128), (65536, 128, 256), (131072, 256, 128),
                  (131072, 128, 128), (131072, 64, 64), (8192, 384, 256), (2048, 832, 512)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda")
    stats = torch.zeros(dense.num_stat_blocks(M), N, 2, device="cuda")
    colmm = torch.zeros(dense.num_stat_blocks(M), N, 2, device="cuda")
    res = []
    for mode in (2, 32, 4, 0):
        lib().p2pb_debug_set(mode)
        res.append(timeit(lambda: dense.gemm_rows([A], W, bias, out=out, stats=stats)))
    lib().p2pb_debug_set(0)
    res.append(timeit(lambda: dense.gemm_rows([A], W, bias, stats=stats, colmm=colmm, store=False)))
    extra = []
    extra.append(timeit(lambda: dense.gemm_rows([A], W, bias, out=out)))            # no stats
    lib().p2pb_debug_set(16)
    extra.append(timeit(lambda: dense.gemm_rows([A], W, bias, out=out, stats=stats)))  # stats combine only, no column loop
    lib().p2pb_debug_set(8)
    extra.append(timeit(lambda: dense.gemm_rows([A], W, bias, out=out, stats=stats)))  # tmem_ld only
    lib().p2pb_debug_set(0)
    extra.append(timeit(lambda: dense.gemm_rows([A], W, None, out=out)))            # no stats, no bias
    print("   experiments: no-stats %.1f | no-col-loop %.1f | tmem_ld only %.1f | no-stats-no-bias %.1f" % tuple(e * 1e3 for e in extra))
    torch.backends.cuda.matmul.allow_tf32 = True
    tl = timeit(lambda: torch.nn.functional.linear(A, W, bias))
    fl = 2.0 * M * K * N; by = 4.0 * (M * K + M * N)
    print(f"M={M:8d} K={K:4d} N={N:4d}: {res[0]*1e3:7.1f} | {res[1]*1e3:7.1f} | {res[2]*1e3:7.1f} | pair {res[3]*1e3:7.1f} | {res[4]*1e3:7.1f} | torch-tf32 {tl*1e3:7.1f}"
          f" | ideal hbm {by/6.4e12*1e6:6.1f} tensor {fl/1.15e15*1e6:6.1f} | best {fl/min(res[:4])/1e9:7.1f} TFLOP/s {by/min(res[:4])/1e6:7.1f} GB/s")

print("conv3d r=8 (per-tap implicit GEMM): us legacy | persist | persist+cluster")
for (B, r, cin, cout) in [(64, 8, 256, 256), (64, 8, 256, 128), (64, 8, 128, 128)]:
    grid = torch.randn(B, r, r, r, cin, device="cuda")
    w = torch.randn(cout, 27 * cin, device="cuda") / (27 * cin) ** 0.5
    bias = torch.randn(cout, device="cuda"); out = torch.empty(B * r ** 3, cout, device="cuda")
    stats = torch.zeros(B * r ** 3 // 32, cout, 2, device="cuda")
    res = []
    for mode in (2, 32, 4, 0):
        lib().p2pb_debug_set(mode)
        res.append(timeit(lambda: dense.conv3d_cl(grid, w, bias, B, r, cin, cout, out=out, stats=stats)))
    lib().p2pb_debug_set(0)
    fl = 2.0 * B * r ** 3 * 27 * cin * cout
    print(f"B={B} r={r} {cin}->{cout}: {res[0]*1e3:7.1f} | {res[1]*1e3:7.1f} | {res[2]*1e3:7.1f} | pair {res[3]*1e3:7.1f} | best {fl/min(res)/1e9:7.1f} TFLOP/s")
