"""Dev tool: BASELINE config 3 (PVDL, 32 patches of 8192 points, xyz only, T = 30) on the engine with its CUDA graph -> patches/s.
usage: python tools/bench_pvdl.py [reps]      (P2PB_LIB=... selects another build of the library, see tools/gpu_ab.sh)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, yaml
import bench
from p2pb_b200.config import Config
from p2pb_b200.model_loader import seeded_state_dict
from p2pb_b200.p2pb import P2PB
from p2pb_b200.unet_pvc import PVCNN2Unet
from tests.helpers import patch_input

dev = torch.device("cuda:0")
cd = yaml.safe_load(open(os.path.join(bench.ROOT, "p2pb_b200", "configs", "PVDL_SNPP.yaml")))
cd["data"]["npoints"] = bench.PVDL_N; cd["model"]["extra_feature_channels"] = 0
cfg = Config.wrap(cd); cfg.gpu = str(dev); cfg.model.ema = False
net = PVCNN2Unet(cfg); net.load_state_dict(seeded_state_dict(net, seed=0), strict=True)
model = P2PB(cfg, net.to(dev)).eval()
x = patch_input(bench.PVDL_B_PER_GPU, bench.PVDL_N, seed=7).to(dev)
run = lambda: model.sample(x_start=x, steps=bench.TSTEPS, log_count=1, verbose=False, use_ema=False)["x_pred"]
for _ in range(2): run()
torch.cuda.synchronize()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"PVDL N={bench.PVDL_N} B={bench.PVDL_B_PER_GPU} T={bench.TSTEPS}: {ms:.1f} ms per call, {bench.PVDL_B_PER_GPU / ms * 1e3:.2f} patches/s, "
      f"{ms / bench.TSTEPS * 1e3:.0f} us per evaluation")
