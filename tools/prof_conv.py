"""Dev tool: run one halo conv + one plain conv (64->64 @ 32^3, B=64) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
B, r, cin, cout = 64, 32, 64, 64
grid = torch.randn(B, r, r, r, cin, device="cuda")
w = torch.randn(cout, cin, 3, 3, 3, device="cuda") / (27 * cin) ** 0.5
wp = dense.pack_conv3d_weight(w, cin)
bias = torch.randn(cout, device="cuda")
out = torch.empty(B * r ** 3, cout, device="cuda")
X = dense.dense_to_padded(grid, r)
_, _, tps = dense.halo_layout(r)
hst = torch.zeros(B * tps, cout, 2, device="cuda")
st = torch.zeros(B * r ** 3 // 32, cout, 2, device="cuda")
for _ in range(3):
    dense.conv3d_halo(X, wp, bias, B, r, cin, cout, out=out, stats=hst)
    dense.conv3d_cl(grid, wp, bias, B, r, cin, cout, out=out, stats=st)
torch.cuda.synchronize()
