"""Dev tool: halo-conv pipeline sweep on the PVDS layer shapes (B=64): (w_stages, a_stages, G) overrides vs automatic;
G < 0 selects the un-paired (cta_group::1) kernel: -1 = automatic G, -k = k tiles per unit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
from p2pb_b200._lib import lib

B = 64


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for r, cin, cin_valid, cout in [(32, 64, 35, 32), (32, 32, 32, 32), (16, 128, 128, 64), (16, 64, 64, 64), (16, 128, 128, 128), (32, 64, 64, 64)]:
    grid = torch.randn(B, r, r, r, cin, device="cuda")
    grid[..., cin_valid:] = 0
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") / (27 * cin) ** 0.5
    wp = dense.pack_conv3d_weight(w, cin)
    bias = torch.randn(cout, device="cuda")
    out = torch.empty(B * r ** 3, cout, device="cuda")
    X = dense.dense_to_padded(grid, r)
    _, _, tps = dense.halo_layout(r)
    hst = torch.zeros(B * tps, cout, 2, device="cuda")
    fl = 2.0 * B * r ** 3 * 27 * cin_valid * cout
    line = f"r={r} {cin}({cin_valid})->{cout}:"
    c64 = (cin + 63) // 64 * 64
    gh = torch.zeros(B, r, r, r, c64, device="cuda", dtype=torch.float16); gh[..., :cin] = grid.half()
    Xh = dense.dense_to_padded(gh, r)
    wh = torch.zeros(cout, c64, 3, 3, 3, device="cuda"); wh[:, :cin] = w
    wph = dense.pack_conv3d_weight(wh, c64).half()
    for cfg in [(0, 0, 0), (0, 0, -1), (0, 0, 2), (0, 0, 3)]:
        lib().p2pb_conv_halo_tune(*cfg)
        try:
            t = timeit(lambda: dense.conv3d_halo(X, wp, bias, B, r, cin, cout, out=out, stats=hst, cin_valid=cin_valid))
            th = timeit(lambda: dense.conv3d_halo(Xh, wph, bias, B, r, c64, cout, out=out, stats=hst, cin_valid=cin_valid))
            line += f"  {cfg}: tf32 {t*1e3:5.0f}us {fl/t/1e9:4.0f}TF | half {th*1e3:5.0f}us {fl/th/1e9:4.0f}TF"
        except Exception as ex:
            line += f"  {cfg}: n/a {ex}"
    lib().p2pb_conv_halo_tune(0, 0, 0)
    print(line, flush=True)
