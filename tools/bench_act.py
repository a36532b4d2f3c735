"""Dev tool: bandwidth of the activation passes (GroupNorm-apply + Swish) at the PVDS bench shapes."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
from p2pb_b200._lib import call, lib
vp = ctypes.c_void_p
p = lambda t: vp(t.data_ptr())
s = lambda: vp(torch.cuda.current_stream().cuda_stream)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B, r, C = 64, 32, 64
raw = torch.randn(B * r ** 3, C, device="cuda"); A = torch.randn(B, C, device="cuda"); Bc = torch.randn(B, C, device="cuda")
outp = dense.alloc_padded(B, C, r, "cuda", torch.float16)
M = 64 * 2048 * 8
x = torch.randn(M, 64, device="cuda"); y = torch.empty(M, 64, device="cuda", dtype=torch.float16)
ref1 = ref2 = None
for g in (0, 4, 8, 16, 32):
    lib().p2pb_set_act_grid(g)
    t1 = timeit(lambda: call("p2pb_affine_act_padded_f16", p(raw), C, p(A), p(Bc), B, C, r, p(outp), C, s()))
    t2 = timeit(lambda: call("p2pb_affine_act_f16", p(x), 64, p(A), p(Bc), M // 64, M, 64, 1, p(y), 64, s()))
    if ref1 is None: ref1, ref2 = outp.clone(), y.clone()
    assert torch.equal(ref1, outp) and torch.equal(ref2, y), "variant differs"
    b1 = raw.numel() * 4 + B * r ** 3 * C * 2; b2 = x.numel() * 6
    print(f"grid {g:2d} CTAs/SM: padded {t1*1e3:6.1f} us {b1/t1/1e6:6.0f} GB/s | rows {t2*1e3:6.1f} us {b2/t2/1e6:6.0f} GB/s")
lib().p2pb_set_act_grid(0)
