import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import load_cfg
from tests.test_model_gpu import build
from p2pb_b200.engine import get_engine, _p, _s, call
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
E = eng.E
sin = eng.buf("temb.sin", T, E); coef = eng.buf("coef", T, 3)
sin.copy_(torch.stack([eng.time_embedding(float(model.noise_levels[s].item()), None) for _, s in pairs]))
coef.copy_(torch.tensor([model.posterior_coefs(p, s) for p, s in pairs]).cuda())
xt = eng.buf("xt", B, 3, N); xt.copy_(x)
xt_e = x.clone()
with torch.no_grad():
    for s, (prev, step) in enumerate(pairs):
        th = eng.buf("temb.h", B, E); temb = eng.buf("temb", B, E)
        row = sin[s:s + 1].expand(B, E)
        eng.linear(row, eng.W["tw0"], eng.W["tb0"], 4, th)
        eng.linear(th, eng.W["tw2"], eng.W["tb2"], 0, temb)
        nl = model.noise_levels[torch.full((B,), step, device="cuda", dtype=torch.long)]
        temb_ref = model.model.embedf(model.model.get_timestep_embedding(nl, "cuda"))
        print(s, "temb diff", float((temb - temb_ref).abs().max()))
        xin = xt.clone()
        eps = eng.evaluate(xt, temb)
        eps_g = eps[:, :3].reshape(B, N, 3).permute(0, 2, 1).clone()
        out_e = model.model(xin, nl, x_cond=None)
        print("   eps diff (same xt)", float((out_e - eps_g).abs().max()), "xt unchanged by evaluate:", bool(torch.equal(xin, xt)))
        call("p2pb_bridge_update", _p(xt), _p(eps), 16, _p(coef[s]), 0, _p(xt), _p(None), B, N, _s())
        st = torch.full((B,), step, device="cuda", dtype=torch.long)
        p0 = model.compute_pred_x0_from_eps(st, xin, eps_g)
        ref_next = model.p_posterior(prev, step, xin, p0)
        print("   bridge update diff", float((ref_next - xt).abs().max()))
        xt_e = ref_next
