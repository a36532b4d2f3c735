"""Dev tool: what the side-stream geometry kernels (FPS 2048 -> 512 of 64 patches: one CTA per patch, 270 us, latency-bound) cost the
persistent one-CTA-per-SM GEMMs that run next to them on the main stream.  Times the three HBM-bound global-PointNet GEMMs alone and
with FPS running concurrently on another stream.
usage: python tools/bench_overlap.py            (P2PB_SMEM_KB=199 leaves room for the FPS CTAs beside a GEMM CTA)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense, ops
from p2pb_b200._lib import lib
if len(sys.argv) > 1: lib().p2pb_gemm_tune(int(sys.argv[1]))      # 32: independent CTAs instead of CTA pairs

B, N = (int(sys.argv[2]) if len(sys.argv) > 2 else 64), 2048
xyz = torch.randn(B, 3, N, device="cuda")
side = torch.cuda.Stream()
shapes = [(131072, 64, 128), (131072, 128, 256), (131072, 256, 512)]
bufs = []
for (M, K, Nn) in shapes:
    A = torch.randn(M, K, device="cuda").half(); W = (torch.randn(Nn, K, device="cuda") / K ** 0.5).half()
    bufs.append((A, W, torch.randn(Nn, device="cuda"), torch.empty(M, Nn, device="cuda"),
                 torch.zeros(dense.num_stat_blocks(M), Nn, 2, device="cuda")))


def gemms():
    for A, W, b, out, st in bufs:
        dense.gemm_rows([A], W, b, out=out, stats=st)


def timed(with_fps, n=10):
    for _ in range(2): gemms()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if with_fps:
            with torch.cuda.stream(side):
                ops.furthest_point_sampling(xyz, 512)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(bufs) + 1)]
        ev[0].record()
        for i, (A, W, b, out, st) in enumerate(bufs):
            dense.gemm_rows([A], W, b, out=out, stats=st); ev[i + 1].record()
        torch.cuda.synchronize()
        ts.append([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(len(bufs))])
    ts.sort(key=sum)
    return "+".join(f"{t:.0f}" for t in ts[len(ts) // 2]) + f" = {sum(ts[len(ts) // 2]):.0f}"


f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.furthest_point_sampling(xyz, 512); torch.cuda.synchronize()
f0.record(); ops.furthest_point_sampling(xyz, 512); f1.record(); torch.cuda.synchronize()
print(f"FPS alone {f0.elapsed_time(f1) * 1e3:.0f} us; 3 GEMMs alone {timed(False)} us; with FPS on another stream {timed(True)} us "
      f"(P2PB_SMEM_KB={os.environ.get('P2PB_SMEM_KB', 'default')})")
