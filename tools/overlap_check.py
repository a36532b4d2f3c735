"""Dev tool: do small kernels run next to a persistent tensor-core kernel of another stream?"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
from p2pb_b200._lib import call, lib

vp = ctypes.c_void_p
p = lambda t: vp(t.data_ptr()) if t is not None else vp(0)
B, r, cin, cout = 32, 32, 64, 64
grid = torch.randn(B, r, r, r, cin, device="cuda")
wp = dense.pack_conv3d_weight(torch.randn(cout, cin, 3, 3, 3, device="cuda") / 40, cin)
bias = torch.randn(cout, device="cuda"); out = torch.empty(B * r ** 3, cout, device="cuda")
X = dense.dense_to_padded(grid, r); _, _, tps = dense.halo_layout(r, cout, False)
hst = torch.zeros(B * tps, cout, 2, device="cuda")
M, C = 32 * 2048 * 8, 64
x = torch.randn(M, C, device="cuda"); y = torch.empty_like(x)
A = torch.randn(32, C, device="cuda"); Bc = torch.randn(32, C, device="cuda")
stats = torch.randn(32 * 64, C, 2, device="cuda"); gam = torch.ones(C, device="cuda")
cA = torch.empty(32, C, device="cuda"); cB = torch.empty(32, C, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def conv(): dense.conv3d_halo(X, wp, bias, B, r, cin, cout, out=out, stats=hst)
def act():
    call("p2pb_affine_act", p(x), C, p(A), p(Bc), M // 32, M, C, 1, 1, p(y), C, vp(0), vp(torch.cuda.current_stream().cuda_stream))
def coef():
    call("p2pb_gn_coef", p(stats), 64, 32, C, 8, 2048, p(gam), p(gam), vp(0), 0, 0, ctypes.c_float(1e-5), p(cA), p(cB), vp(0),
         vp(torch.cuda.current_stream().cuda_stream))

def timed(fa, na, fb, nb):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        for _ in range(na): fa()
    with torch.cuda.stream(s2):
        for _ in range(nb): fb()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

nop = lambda: None
for kb in (227, 190):
    lib().p2pb_set_smem_budget_kb(kb)
    for name, f, n in (("affine_act 134MB", act, 20), ("gn_coef", coef, 200)):
        for _ in range(2):
            tc = timed(conv, 10, nop, 0); ts = timed(nop, 0, f, n); tb = timed(conv, 10, f, n)
        print(f"smem {kb} KB: 10 x conv {tc:.2f} ms | {n} x {name} {ts:.2f} ms | both streams {tb:.2f} ms (sum {tc+ts:.2f})")
