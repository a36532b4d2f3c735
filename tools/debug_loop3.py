import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import load_cfg
from tests.test_model_gpu import build
z = np.load("tests/golden/model_pvds_cfg1.npz")
cfg = load_cfg("PVDS_PUNet")
x = torch.from_numpy(z["x_start"]).cuda()
for mode in ("nograph", "graph"):
    if mode == "nograph": os.environ["P2PB_NO_GRAPH"] = "1"
    else: os.environ.pop("P2PB_NO_GRAPH", None)
    model, _ = build(cfg, backend="engine")
    out = model.sample(x_start=x, steps=5, log_count=5, verbose=False)
    for i in range(5):
        print(mode, "chain", i, float((out["x_chain"][:, i].cpu() - torch.from_numpy(z["x_chain"][:, i])).abs().max()))
