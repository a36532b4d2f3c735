"""Dev tool: time the tcgen05 conv / GEMM kernels on the PVDS layer shapes (B=64) next to cuDNN (TF32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from p2pb_b200 import dense

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
shapes = [(32, 64, 32), (32, 32, 32), (16, 128, 64), (16, 64, 64), (8, 192, 128), (8, 128, 128), (8, 256, 256), (16, 128, 128), (32, 64, 64)]

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

torch.backends.cudnn.benchmark = True
tot_m = tot_c = 0
for r, cin, cout in shapes:
    grid = torch.randn(B, r, r, r, cin, device="cuda")
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") / (27 * cin) ** 0.5
    wp = dense.pack_conv3d_weight(w, cin)
    bias = torch.randn(cout, device="cuda")
    out = torch.empty(B * r ** 3, cout, device="cuda")
    stats = torch.zeros(B * r ** 3 // 32, cout, 2, device="cuda")
    t = timeit(lambda: dense.conv3d_cl(grid, wp, bias, B, r, cin, cout, out=out, stats=stats))
    x = grid.permute(0, 4, 1, 2, 3)  # NCDHW view with channels_last_3d strides
    tc = timeit(lambda: F.conv3d(x, w, bias, padding=1))
    xc = x.contiguous()
    tcc = timeit(lambda: F.conv3d(xc, w, bias, padding=1))
    th = float("nan")
    if r >= 16 and cout <= 128:
        X = dense.dense_to_padded(grid, r)
        _, _, tps = dense.halo_layout(r, cout, False)
        hst = torch.zeros(B * tps, cout, 2, device="cuda")
        th = timeit(lambda: dense.conv3d_halo(X, wp, bias, B, r, cin, cout, out=out, stats=hst))
    fl = 2.0 * B * r ** 3 * 27 * cin * cout
    tot_m += t; tot_c += min(tc, tcc)
    print(f"r={r:3d} cin={cin:4d} cout={cout:4d}: ours {t:7.3f} ms {fl / t / 1e9:7.1f} TF/s | halo {th:7.3f} ms {fl / th / 1e9:7.1f} | cudnn CL {tc:7.3f} ms {fl / tc / 1e9:7.1f} | cudnn NCDHW {tcc:7.3f} ms {fl / tcc / 1e9:7.1f}")
print("sum ours", tot_m, "sum cudnn best", tot_c)
