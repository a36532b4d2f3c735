"""Dev tool: where the time of an epilogue-bound GEMM goes -- the same launch with parts of the epilogue switched off through the
arguments (no output, no statistics, no bias).  usage: python tools/bench_gemm_parts.py [M K N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense

M, K, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (131072, 128, 256)
R = 6
As = [torch.randn(M, K, device="cuda").half() for _ in range(R)]
W = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
bias = torch.randn(N, device="cuda"); outs = [torch.empty(M, N, device="cuda") for _ in range(R)]
stats = torch.zeros(dense.num_stat_blocks(M), N, 2, device="cuda"); colmm = torch.zeros_like(stats)


def timeit(fn, n=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, kw in [("store + stats + bias", dict(bias=bias, stats=stats)), ("store + bias", dict(bias=bias)), ("store only", dict()),
                 ("stats + colmm + bias, no store", dict(bias=bias, stats=stats, colmm=colmm, store=False)),
                 ("stats + bias, no store", dict(bias=bias, stats=stats, store=False)),
                 ("colmm only, no bias, no store", dict(colmm=colmm, store=False))]:
    t = timeit(lambda i: dense.gemm_rows([As[i % R]], W, out=outs[i % R] if kw.get("store", True) else None, **kw))
    print(f"{M}x{K}->{N}  {name:40s} {t:7.1f} us")
