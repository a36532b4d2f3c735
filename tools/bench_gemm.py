"""Dev tool: time rows-GEMMs of the PVDS evaluation (B=64) with/without stats and stores."""
import sys, os, ctypes
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

for (M, K, N) in [(131072, 128, 256), (131072, 512, 1024), (131072, 32, 128), (1048576, 64, 32), (1048576, 32, 64), (262144, 96, 64)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda")
    stats = torch.zeros(dense.num_m_tiles(M), N, 2, device="cuda")
    res = []
    for dbg, st in [(0, stats), (0, None), (1, stats), (1, None)]:
        lib().p2pb_debug_set(dbg)
        res.append(timeit(lambda: dense.gemm_rows([A], W, bias, out=out, stats=st)))
    lib().p2pb_debug_set(0)
    tl = timeit(lambda: torch.nn.functional.linear(A, W, bias))
    fl = 2.0 * M * K * N; by = 4.0 * (M * K + M * N)
    print(f"M={M} K={K} N={N}: full {res[0]*1e3:7.1f}us | no-stats {res[1]*1e3:7.1f} | no-store {res[2]*1e3:7.1f} | neither {res[3]*1e3:7.1f} | torch {tl*1e3:7.1f}us | ideal hbm {by/6.4e12*1e6:6.1f}us tensor {fl/1.15e15*1e6:6.1f}us")
