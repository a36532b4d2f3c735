"""Dev tool: halo-conv A/B on the PVDS / PVDL layer shapes (64 patches, half operands): dz-stacked form (N = 3 Cout per MMA) vs the
un-stacked round-1 kernel (p2pb_conv_halo_tune(0, 0, 100)), and tiles-per-unit variants.  CUDA events, 10 launches each
(every launch moves >> L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
from p2pb_b200._lib import lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for r, cin, cin_valid, cout in [(32, 64, 35, 32), (32, 64, 32, 32), (16, 192, 192, 64), (16, 64, 64, 64), (32, 64, 64, 64), (16, 128, 128, 128),
                                (32, 128, 67, 64), (32, 128, 128, 128)]:
    gh = torch.zeros(B, r, r, r, cin, device="cuda", dtype=torch.float16)
    gh[..., :cin_valid] = torch.randn(B, r, r, r, cin_valid, device="cuda").half()
    Xh = dense.dense_to_padded(gh, r)
    del gh
    wph = (torch.randn(cout, 27 * cin, device="cuda") / (27 * cin) ** 0.5).half()
    bias = torch.randn(cout, device="cuda")
    out = torch.empty(B * r ** 3, cout, device="cuda")
    hst = torch.zeros(B * 300 * (r // 16) ** 3, cout, 2, device="cuda")
    fl = 2.0 * B * r ** 3 * 27 * cin_valid * cout
    line = f"r={r} {cin}({cin_valid})->{cout} B={B}:"
    for name, cfg in [("stacked", (0, 0, 0)), ("stacked G=1", (0, 0, 1)), ("stacked G=2", (0, 0, 2)), ("un-stacked", (0, 0, 100)),
                      ("stacked, no pairs", (0, 0, -1))]:
        lib().p2pb_conv_halo_tune(*cfg)
        try:
            th = timeit(lambda: dense.conv3d_halo(Xh, wph, bias, B, r, cin, cout, out=out, stats=hst, cin_valid=cin_valid))
            line += f"  {name}: {th*1e3:5.0f} us {fl/th/1e9:4.0f} TF"
        except Exception as ex:
            line += f"  {name}: n/a ({str(ex)[:60]})"
    lib().p2pb_conv_halo_tune(0, 0, 0)
    print(line, flush=True)
    del Xh, out, hst
