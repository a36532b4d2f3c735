"""Dev tool: PVDL (configs 3-4 of BASELINE.json: N=8192, data.npoints=8192) throughput of the engine, T steps.
usage: python tools/bench_pvdl.py [batch=32] [T=30] [extra_channels=0] [nograph]     (P2PB_LIB=... selects another build of the library,
see tools/gpu_ab_pvdl.sh)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, yaml
from p2pb_b200.config import Config
from p2pb_b200.model_loader import seeded_state_dict
from p2pb_b200.p2pb import P2PB
from p2pb_b200.unet_pvc import PVCNN2Unet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
extra = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N = 8192
if len(sys.argv) > 4 and sys.argv[4] == "nograph":      # ncu launch lists cannot see inside a CUDA graph
    from p2pb_b200 import engine as _E
    _E.OPTIONS.no_graph = True
cfg_dict = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "..", "p2pb_b200", "configs", "PVDL_SNPP.yaml")))
cfg_dict["data"]["npoints"] = N
cfg_dict["model"]["extra_feature_channels"] = extra
cfg = Config.wrap(cfg_dict); cfg.gpu = "cuda:0"; cfg.model.ema = False
net = PVCNN2Unet(cfg); net.load_state_dict(seeded_state_dict(net, seed=0))
model = P2PB(cfg, net.cuda()).eval()
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 3, N, generator=g); x = x / x.norm(dim=1, keepdim=True) + 0.02 * torch.randn(B, 3, N, generator=g)
x = x - x.mean(2, keepdim=True); x = (x / x.norm(dim=1).amax(dim=1)[:, None, None]).contiguous().cuda()
xc = torch.rand(B, extra, N, generator=g).cuda() if extra else None
for _ in range(2):
    model.sample(x_start=x, x_cond=xc, steps=T, log_count=1, verbose=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 2
for _ in range(n):
    model.sample(x_start=x, x_cond=xc, steps=T, log_count=1, verbose=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"PVDL N={N} extra={extra} B={B} T={T}: {ms:.1f} ms per sample() -> {B / ms * 1e3:.2f} patches/s, {ms / T:.2f} ms per evaluation; "
      f"algorithmic 333 GFLOP/patch/step -> {333e9 * B * T / ms / 1e9:.0f} TFLOP/s")
