"""Dev tool: run the halo conv as the engine launches it (half operands, CTA pairs, B = 64 = the bench batch) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
B = 64
for r, cin, cout in [(32, 64, 64), (16, 128, 128)]:
    grid = torch.randn(B, r, r, r, cin, device="cuda").half()
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") / (27 * cin) ** 0.5
    wp = dense.pack_conv3d_weight(w, cin).half()
    bias = torch.randn(cout, device="cuda")
    out = torch.empty(B * r ** 3, cout, device="cuda")
    X = dense.dense_to_padded(grid, r)
    _, _, tps = dense.halo_layout(r)
    hst = torch.zeros(B * tps, cout, 2, device="cuda")
    for _ in range(4):
        dense.conv3d_halo(X, wp, bias, B, r, cin, cout, out=out, stats=hst)
    torch.cuda.synchronize()
