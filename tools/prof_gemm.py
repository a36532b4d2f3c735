"""Dev tool (run under ncu): one HBM-bound rows-GEMM of the evaluation with half operands, GroupNorm partials, fp32 output."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from p2pb_b200 import dense
M, K, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (262144, 64, 128)
A = torch.randn(M, K, device="cuda").half(); W = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda")
stats = torch.zeros(dense.num_stat_blocks(M), N, 2, device="cuda")
for _ in range(4):
    dense.gemm_rows([A], W, bias, out=out, stats=stats)
torch.cuda.synchronize()
