import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, yaml
from tests.helpers import load_cfg
from tests.test_model_gpu import build
from p2pb_b200.engine import get_engine
from p2pb_b200.p2pb import space_indices

torch.backends.cudnn.allow_tf32 = False
z = np.load("tests/golden/model_pvds_cfg1.npz")
cfg = load_cfg("PVDS_PUNet")
model, _ = build(cfg, backend="engine")
x = torch.from_numpy(z["x_start"]).cuda()
B, _, N = x.shape
eng = get_engine(model, model.model, x.shape, None)
T = 5
steps = space_indices(1000, T + 1)
rev = steps[::-1]
pairs = list(zip(rev[1:], rev[:-1]))
xt_e = x.clone()   # eager trajectory
xt_g = x.clone()   # engine trajectory
with torch.no_grad():
    for prev, step in pairs:
        nl = model.noise_levels[torch.full((B,), step, device="cuda", dtype=torch.long)]
        out_e = model.model(xt_e, nl, x_cond=None)
        # engine eval on the SAME xt as eager
        sin = eng.time_embedding(float(nl[0].item()), None)[None].expand(B, -1).contiguous()
        th, temb = eng.buf("t.th", B, eng.E), eng.buf("t.temb", B, eng.E)
        eng.linear(sin, eng.W["tw0"], eng.W["tb0"], 4, th)
        eng.linear(th, eng.W["tw2"], eng.W["tb2"], 0, temb)
        eps_rows = eng.evaluate(xt_e.contiguous(), temb)
        eps_g = eps_rows[:, :3].reshape(B, N, 3).permute(0, 2, 1)
        print(f"step {step}->{prev}: nl={nl[0].item():.3f} |eps_eager-eps_engine| mean={float((out_e-eps_g).abs().mean()):.3e} max={float((out_e-eps_g).abs().max()):.3e}")
        st = torch.full((B,), step, device="cuda", dtype=torch.long)
        p0 = model.compute_pred_x0_from_eps(st, xt_e, out_e)
        xt_e = model.p_posterior(prev, step, xt_e, p0)
        c = model.posterior_coefs(prev, step)
        print("   coefs", c, "std_fwd", float(model.std_fwd[step]))
out = model.sample(x_start=x, steps=T, log_count=T, verbose=False)
print("final eager vs golden", float((xt_e.cpu() - torch.from_numpy(z["x_pred"])).abs().max()))
print("final engine.sample vs eager", float((out["x_pred"] - xt_e).abs().max()))
for i in range(T):
    print(" chain", i, float((out["x_chain"][:, i].cpu() - torch.from_numpy(z["x_chain"][:, i])).abs().max()))
