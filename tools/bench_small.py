"""Dev tool: the small fused per-step kernels at the PVDS bench shapes (64 patches), CUDA events, L2 flushed between launches."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200._lib import call

vp = ctypes.c_void_p
p = lambda t: vp(t.data_ptr()) if t is not None else vp(0)
s = vp(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(64 * 1024 * 1024, device="cuda")


def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


B, E = 64, 64
rn = lambda *sh: torch.randn(*sh, device="cuda")
sin, w0, b0, w2, b2 = rn(1, E), rn(E, E), rn(E), rn(E, E), rn(E)
folds = [rn(c, E) for c in (64, 128, 256, 256, 256, 128, 64, 64)]
bufs = [torch.zeros(B, w.shape[0], device="cuda") for w in folds]
ptr = torch.tensor([b.data_ptr() + 4 * o for b, w in zip(bufs, folds) for o in range(w.shape[0])], dtype=torch.int64, device="cuda")
stride = torch.tensor([w.shape[0] for w in folds for _ in range(w.shape[0])], dtype=torch.int32, device="cuda")
wall = torch.cat(folds, 0).contiguous(); temb = torch.zeros(B, E, device="cuda")
print("step_vectors R=%d: %.1f us" % (wall.shape[0], timeit(lambda: call("p2pb_step_vectors", p(sin), 0, p(w0), p(b0), p(w2), p(b2), B, E, p(wall), wall.shape[0], p(ptr), p(stride), p(temb), s))))
for C in (32, 64, 128, 256):
    ym, v0, v2, se = rn(B, C), rn(C // 8, C), rn(C, C // 8), torch.zeros(B, C, device="cuda")
    print("se_excite C=%d: %.1f us" % (C, timeit(lambda: call("p2pb_se_excite", p(ym), p(v0), p(v2), B, C, C // 8, p(se), s))))
N, C = 2048, 128
raw, A, Bc, W, bias = rn(B * N, C), rn(B, C), rn(B, C), rn(3, C), rn(3)
xt, coef, eps = rn(B, 3, N), torch.tensor([0.7, 0.3, 0.65], device="cuda"), torch.zeros(B * N, 16, device="cuda")
t = timeit(lambda: call("p2pb_head_bridge", p(raw), C, p(A), p(Bc), p(W), p(bias), B, C, N, p(xt), p(coef), 0, p(xt), vp(0), p(eps), 16, s))
print("head_bridge %d points x %d ch: %.1f us = %.0f GB/s" % (B * N, C, t, raw.numel() * 4 / t / 1e3))
