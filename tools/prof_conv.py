"""Dev tool: run the halo conv as the engine launches it (half operands, CTA pairs, B = 64 = the bench batch) for ncu.
usage: python tools/prof_conv.py [tune_G]   (tune_G = third argument of p2pb_conv_halo_tune: 0 default, 100 = un-stacked round-1 form)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
from p2pb_b200._lib import lib
B = 64
G = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib().p2pb_conv_halo_tune(0, 0, G)
shapes = [(32, 64, 64), (32, 64, 32), (16, 128, 128)] if len(sys.argv) < 3 else [(32, 64, 64)]
for r, cin, cout in shapes:
    grid = torch.randn(B, r, r, r, cin, device="cuda").half()
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") / (27 * cin) ** 0.5
    wp = dense.pack_conv3d_weight(w, cin).half()
    bias = torch.randn(cout, device="cuda")
    out = torch.empty(B * r ** 3, cout, device="cuda")
    X = dense.dense_to_padded(grid, r)
    hst = torch.zeros(B * 300 * (r // 16) ** 3, cout, 2, device="cuda")
    for _ in range(3):
        dense.conv3d_halo(X, wp, bias, B, r, cin, cout, out=out, stats=hst)
    torch.cuda.synchronize()
