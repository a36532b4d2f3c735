"""CPU restatement of the reference's denoising hot path -- TEST INFRASTRUCTURE ONLY.

State-dict driven, functional restatement (plain torch fp32 on the host + ``oracle/ops.py``) of

  * the bridge schedule and the T-step sampling loop   ``/root/reference/models/p2pb.py``
  * the PVCNN U-Net forward                            ``/root/reference/models/unet_pvc.py``, ``models/pvcnn.py``,
                                                        ``models/modules.py``

It is the checker for the CUDA path (tests/, smoke(), bench.py cpu_baseline) and never part of the product.
It is pinned by ``oracle/gen_golden.py``, which imports the reference's REAL ``models/*.py`` in the build
container, runs both on identical seeded weights/inputs and commits the outputs to ``tests/golden/``.

``cfg`` is the plain nested dict ``yaml.safe_load`` gives for ``opt.yaml``; ``sd`` is the backbone state dict
(``PVCNN2Unet.state_dict()`` key names, e.g. ``sa_layers.0.0.voxel_layers.0.weight``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import ops as O


# ----------------------------------------------------------------------------------------------------------
# schedule  (models/p2pb.py:16-40, 54-67, 93-130)
# ----------------------------------------------------------------------------------------------------------
def space_indices(num_steps: int, count: int) -> List[int]:
    """p2pb.py:16-40 -- python round() (banker's rounding) on an accumulated float stride."""
    assert count <= num_steps
    stride = 1 if count <= 1 else (num_steps - 1) / (count - 1)
    cur, out = 0.0, []
    for _ in range(count):
        out.append(round(cur))
        cur += stride
    return out


def build_schedule(cfg: dict) -> Dict[str, torch.Tensor]:
    """p2pb.py:93-130: float64 numpy tables cast to fp32."""
    d = cfg["diffusion"]
    n = d["timesteps"]
    scale = 1000 / n
    lo, hi = d["beta_start"] * scale, d["beta_end"] * scale
    betas = (torch.linspace(lo ** 0.5, hi ** 0.5, n, dtype=torch.float64) ** 2).numpy()  # p2pb.py:62-67
    if d.get("symmetric", True):
        betas = np.concatenate([betas[: n // 2], np.flip(betas[: n // 2])])
    noise_levels = torch.linspace(d["t0"], d["T"], n, dtype=torch.float32) * n
    std_fwd = np.sqrt(np.cumsum(betas))
    std_bwd = np.sqrt(np.flip(np.cumsum(np.flip(betas))))
    denom = std_fwd ** 2 + std_bwd ** 2
    mu_x0, mu_x1, var = std_bwd ** 2 / denom, std_fwd ** 2 / denom, (std_fwd ** 2 * std_bwd ** 2) / denom
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32)
    return {"betas": t(betas), "std_fwd": t(std_fwd), "std_bwd": t(std_bwd), "std_sb": t(np.sqrt(var)),
            "mu_x0": t(mu_x0), "mu_x1": t(mu_x1), "noise_levels": noise_levels}


# ----------------------------------------------------------------------------------------------------------
# small layers (models/modules.py)
# ----------------------------------------------------------------------------------------------------------
def swish(x):  # modules.py:25-27
    return x * torch.sigmoid(x)


def _pw(sd, key, x):
    """1x1 Conv1d/Conv2d (pvcnn.py:174-192) as a channel contraction on [B,C,...]."""
    w = sd[key + ".weight"]
    w2 = w.reshape(w.shape[0], w.shape[1])
    y = torch.einsum("oc,bc...->bo...", w2, x)
    b = sd.get(key + ".bias")
    if b is not None:
        y = y + b.reshape(1, -1, *([1] * (x.dim() - 2)))
    return y


def _gn(x, groups, w, b, eps=1e-5):
    return F.group_norm(x, groups, w, b, eps)


def adagn(sd, key, x, cond, groups=8):
    """modules.py:319-358: GroupNorm(x) * factor + bias with [factor, bias] = Linear(cond)."""
    e = F.linear(cond, sd[key + ".emd.weight"], sd[key + ".emd.bias"])
    e = e.reshape(e.shape[0], -1, *([1] * (x.dim() - 2)))
    factor, bias = e.chunk(2, 1)
    return _gn(x, groups, sd[key + ".norm.weight"], sd[key + ".norm.bias"]) * factor + bias


def norm_layer(sd, key, x, cond, groups=8):
    """SharedMLP/PVConv norm: AdaGN when cond_dim>0 else plain GroupNorm (pvcnn.py:179-182, 260-263)."""
    if key + ".emd.weight" in sd:
        return adagn(sd, key, x, cond, groups)
    return _gn(x, groups, sd[key + ".weight"], sd[key + ".bias"])


def shared_mlp(sd, key, x, cond):
    """pvcnn.py:162-205: (conv1x1 -> norm -> Swish)* ; layer indices 0,1,2 / 3,4,5 / ..."""
    i = 0
    while f"{key}.layers.{i}.weight" in sd:
        x = _pw(sd, f"{key}.layers.{i}", x)
        x = norm_layer(sd, f"{key}.layers.{i + 1}", x, cond)
        x = swish(x)
        i += 3
    return x


def se3d(sd, key, x):
    """modules.py:362-378."""
    m = x.mean(-1).mean(-1).mean(-1)
    s = torch.sigmoid(F.linear(F.relu(F.linear(m, sd[key + ".fc.0.weight"])), sd[key + ".fc.2.weight"]))
    return x * s.reshape(x.shape[0], x.shape[1], 1, 1, 1)


def linear_attention(sd, key, x, heads):
    """modules.py:165-194 (no residual, q neither scaled nor soft-maxed)."""
    B, C, N = x.shape
    qkv = _pw(sd, key + ".to_qkv", x)  # [B, 3*h*32, N]
    q, k, v = qkv.reshape(B, 3, heads, -1, N).unbind(1)
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, -1, N)
    return _pw(sd, key + ".to_out", out)


def softmax_attention(sd, key, x, heads):
    """modules.py:197-264 (Attention with norm=False, no time_cond, no qk_norm) around Attend (:77-162, scaled dot-product softmax),
    with the transposes of unet_pvc.py:239-241: x [B,C,N] -> [B,C,N]."""
    B, C, N = x.shape
    t = x.transpose(1, 2)
    q = F.linear(t, sd[key + ".to_q.weight"])
    k, v = F.linear(t, sd[key + ".to_kv.weight"]).chunk(2, dim=-1)
    q, k, v = (z.reshape(B, N, heads, -1).transpose(1, 2) for z in (q, k, v))
    attn = (torch.einsum("bhid,bhjd->bhij", q, k) * (q.shape[-1] ** -0.5)).softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(1, 2).reshape(B, N, -1)
    return F.linear(out, sd[key + ".to_out.weight"]).transpose(1, 2)


def timestep_embedding(t, dim):
    """unet_pvc.py:156-169."""
    half = dim // 2
    e = np.log(10000) / (half - 1)
    e = torch.from_numpy(np.exp(np.arange(0, half) * -e)).float()
    e = t[:, None] * e[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if dim % 2 == 1:
        e = F.pad(e, (0, 1))
    return e


# ----------------------------------------------------------------------------------------------------------
# blocks (models/pvcnn.py)
# ----------------------------------------------------------------------------------------------------------
def pvconv(sd, key, feats, coords, cond, r):
    """pvcnn.py:306-334 PVConv.forward (eval)."""
    B, C, N = feats.shape
    norm_coords, vox = O.voxel_coords(coords, r)                                   # pvcnn.py:215-231
    grid = O.avg_voxelize_forward(feats, vox, r)[0].reshape(B, C, r, r, r)
    v = F.conv3d(grid, sd[key + ".voxel_layers.0.weight"], sd[key + ".voxel_layers.0.bias"], padding=1)
    v = swish(norm_layer(sd, key + ".voxel_layers.1", v, cond))
    v = F.conv3d(v, sd[key + ".voxel_layers.4.weight"], sd[key + ".voxel_layers.4.bias"], padding=1)
    v = norm_layer(sd, key + ".voxel_layers.5", v, cond)
    if key + ".voxel_layers.6.fc.0.weight" in sd:
        v = se3d(sd, key + ".voxel_layers.6", v)
    vf = O.trilinear_devoxelize_forward(r, False, norm_coords, v.reshape(B, -1, r ** 3))[0]
    return vf + shared_mlp(sd, key + ".point_features", feats, cond)


def sa_module(sd, key, feats, coords, cond, num_centers, radius, K=32):
    """pvcnn.py:388-424 + BallQuery 111-127."""
    idx = O.furthest_point_sampling_forward(coords, num_centers)
    centers = O.gather_features_forward(coords, idx)
    nidx = O.ball_query(centers, coords, radius, K)
    g_xyz = O.grouping_forward(coords, nidx) - centers.unsqueeze(-1)
    g = torch.cat([g_xyz, O.grouping_forward(feats, nidx)], dim=1)
    g = shared_mlp(sd, key + ".mlps.0", g, cond)
    return g.max(dim=-1).values, centers


def fp_module(sd, key, coords, lower_coords, lower_feats, skip, cond):
    """pvcnn.py:446-467."""
    x = O.three_nearest_neighbors_interpolate_forward(coords, lower_coords, lower_feats)[0]
    if skip is not None:
        x = torch.cat([x, skip], dim=1)
    return shared_mlp(sd, key + ".mlp", x, cond)


def global_pnet(sd, key, coords):
    """pvcnn.py:905-932 Pnet2Stage with ConditionedSharedMLPLayer 826-902 (no time/cond embedding, no residual)."""
    def mlp(k, x):  # MLP 803-823: conv(bias) -> MyGroupNorm(32) -> Swish
        x = _pw(sd, k + ".mlp.0", x)
        w, b = sd[k + ".mlp.1.group_norm.weight"], sd[k + ".mlp.1.group_norm.bias"]
        nc = w.shape[0]
        assert nc == x.shape[1]
        return swish(_gn(x, 32, w, b))

    x = coords
    x = mlp(key + ".mlp1.shared_mlp_0", x)
    x = mlp(key + ".mlp1.shared_mlp_1", x)
    g = x.max(dim=2, keepdim=True).values.expand(-1, -1, x.shape[2])
    x = torch.cat([x, g], dim=1)
    x = mlp(key + ".mlp2.shared_mlp_0", x)
    x = mlp(key + ".mlp2.shared_mlp_1", x)
    return x.max(dim=2).values


def model_plan(cfg: dict) -> dict:
    """Static structure from opt.yaml (pvcnn.py:34-96 create_pvc_layer_params, 528-665, 668-741)."""
    m, pvd = cfg["model"], cfg["model"]["PVD"]
    ch = pvd["channels"]
    npoints = cfg["data"]["npoints"]
    vr, rad = pvd["voxel_resolutions"], pvd["radius"]
    nsa, nfp = pvd["n_sa_blocks"], pvd["n_fp_blocks"]
    centers = pvd.get("centers")
    L = len(ch) - 1
    sa = []
    for i in range(L):
        nc = npoints // 4 ** (i + 1) if centers is None else centers[i]
        has_conv = i != L - 1
        # n_sa_blocks only matters at level 0 (pvcnn.py:615-618): deeper levels get exactly one PVConv
        nblk = (nsa[i] if i == 0 else 1) if has_conv else 0
        sa.append({"n_pvconv": nblk, "res": vr[i] if has_conv else None, "centers": nc, "radius": rad[i]})
    fp = [{"n_pvconv": nfp[3], "res": vr[3]}, {"n_pvconv": nfp[2], "res": vr[2]},
          {"n_pvconv": nfp[1], "res": vr[1]}, {"n_pvconv": nfp[0], "res": vr[0]}]
    extra = pvd.get("extra_feature_channels", m.get("extra_feature_channels", 0))
    return {"sa": sa, "fp": fp, "heads": pvd["attention_heads"], "embed_dim": m.get("time_embed_dim") or 64,
            "in_dim": m.get("in_dim") or 3, "extra": extra}


def unet_forward(sd: Dict[str, torch.Tensor], cfg: dict, x, t, x_cond=None, taps: Optional[dict] = None):
    """unet_pvc.py:171-269 PVCNN2Unet.forward (eval mode: Dropout = identity)."""
    plan = model_plan(cfg)
    if x_cond is not None:
        x = torch.cat([x, x_cond], dim=1)
    B, C, N = x.shape
    ind = plan["in_dim"]
    coords = x[:, :ind].contiguous()
    feats = x[:, ind:].contiguous()
    if "embed_feats.0.weight" in sd:                                           # unet_pvc.py:73-83,184-188
        f = coords if plan["extra"] == 0 else feats
        f = _pw(sd, "embed_feats.0", f)
        f = swish(_gn(f, 8, sd["embed_feats.1.weight"], sd["embed_feats.1.bias"]))
        feats = _pw(sd, "embed_feats.3", f)
    cond = global_pnet(sd, "global_pnet", coords) if "global_pnet.mlp1.shared_mlp_0.mlp.0.weight" in sd else None
    feats = torch.cat([coords, feats], dim=1)
    temb = timestep_embedding(t, plan["embed_dim"])
    temb = F.linear(F.leaky_relu(F.linear(temb, sd["embedf.0.weight"], sd["embedf.0.bias"]), 0.1),
                    sd["embedf.2.weight"], sd["embedf.2.bias"])               # [B, 64]
    if taps is not None:
        taps["cond"], taps["temb"] = cond, temb

    def tcat(f):  # cat[features, time_emb.expand(N)]
        return torch.cat([f, temb[:, :, None].expand(-1, -1, f.shape[2])], dim=1)

    skips, coords_list = [], []
    for i, lv in enumerate(plan["sa"]):
        skips.append(feats)
        coords_list.append(coords)
        if i > 0:
            feats = tcat(feats)
        nseq = lv["n_pvconv"] + 1
        for k in range(lv["n_pvconv"]):
            feats = pvconv(sd, f"sa_layers.{i}.{k}", feats, coords, cond, lv["res"])
        key = f"sa_layers.{i}.{lv['n_pvconv']}" if nseq > 1 else f"sa_layers.{i}"
        feats, coords = sa_module(sd, key, feats, coords, cond, lv["centers"], lv["radius"])
        if taps is not None:
            taps[f"sa{i}"] = feats
    if "global_att.to_qkv.weight" in sd:
        feats = linear_attention(sd, "global_att", feats, plan["heads"])
    elif "global_att.to_q.weight" in sd:
        feats = softmax_attention(sd, "global_att", feats, plan["heads"])
    for j, lv in enumerate(plan["fp"]):
        skip, up_coords = skips[-1 - j], coords_list[-1 - j]
        nseq = lv["n_pvconv"] + 1
        key = f"fp_layers.{j}.0" if nseq > 1 else f"fp_layers.{j}"
        feats = fp_module(sd, key, up_coords, coords, tcat(feats), skip, cond)
        coords = up_coords
        for k in range(lv["n_pvconv"]):
            feats = pvconv(sd, f"fp_layers.{j}.{k + 1}", feats, coords, cond, lv["res"])
        if taps is not None:
            taps[f"fp{j}"] = feats
    feats = shared_mlp(sd, "classifier.0", feats, None)
    return _pw(sd, "classifier.2", feats)


# ----------------------------------------------------------------------------------------------------------
# sampling loop (models/p2pb.py:190-363)
# ----------------------------------------------------------------------------------------------------------
def sample(sd, cfg, x_start, x_cond=None, steps=None, log_count=10, clip=False, eps_fn=None):
    """P2PB.sample -> ddpm_sampling -> sample_ddpm with ot_ode (deterministic) posterior."""
    d = cfg["diffusion"]
    assert d.get("sampling_strategy", "DDPM") == "DDPM" and d.get("objective", "pred_noise") == "pred_noise"
    assert d["ot_ode"], "oracle restates the deterministic OT-ODE posterior (all shipped configs)"
    assert not d.get("cond_x1", False) and not d.get("add_x1_noise", False)
    sch = build_schedule(cfg)
    T = d["timesteps"]
    nsteps = steps or d["sampling_timesteps"]
    idx = space_indices(T, nsteps + 1)
    log_count = min(len(idx) - 1, log_count)
    log_steps = [idx[i] for i in space_indices(len(idx) - 1, log_count)]
    rev = idx[::-1]
    xt = x_start.clone()
    xs, x0s = [], []
    if eps_fn is None:
        eps_fn = lambda xt_, nl: unet_forward(sd, cfg, xt_, nl, x_cond)
    B = xt.shape[0]
    for prev, step in zip(rev[1:], rev[:-1]):
        nl = sch["noise_levels"][torch.full((B,), step, dtype=torch.long)]
        out = eps_fn(xt, nl)
        pred_x0 = xt - sch["std_fwd"][step] * out                                # p2pb.py:155-165
        if clip:
            pred_x0 = pred_x0.clamp(-3.0, 3.0)
        std_n, std_p = sch["std_fwd"][step], sch["std_fwd"][prev]               # p2pb.py:190-213
        std_delta = (std_n ** 2 - std_p ** 2).sqrt()
        den = std_p ** 2 + std_delta ** 2
        mu_x0, mu_xn = std_delta ** 2 / den, std_p ** 2 / den
        xt = mu_x0 * pred_x0 + mu_xn * xt
        if prev in log_steps:
            xs.append(xt)
            x0s.append(pred_x0)
    chain = torch.flip(torch.stack(xs, dim=1), dims=(1,))
    return {"x_chain": chain, "x_pred": chain[:, 0], "x_start": x_start,
            "pred_x0": torch.flip(torch.stack(x0s, dim=1), dims=(1,))}


# ----------------------------------------------------------------------------------------------------------
# parameter inventory + seeded synthetic checkpoints (no trained weights exist offline)
# ----------------------------------------------------------------------------------------------------------
def param_shapes(cfg: dict) -> Dict[str, tuple]:
    """Key -> shape of ``PVCNN2Unet(cfg).state_dict()`` (unet_pvc.py:27-154; checked key-by-key against the
    reference's real constructor by oracle/gen_golden.py)."""
    plan = model_plan(cfg)
    m, pvd = cfg["model"], cfg["model"]["PVD"]
    ch = pvd["channels"]
    E = plan["embed_dim"]
    ind, extra = plan["in_dim"], plan["extra"]
    cd = pvd["global_embedding_dim"] if pvd.get("use_global_embedding") else 0
    fe = pvd.get("feat_embed_dim", extra)
    use_se = pvd.get("use_se", True)
    S: Dict[str, tuple] = {}

    def norm(key, c, cond=True):
        if cond and cd > 0:
            S[key + ".norm.weight"] = (c,); S[key + ".norm.bias"] = (c,)
            S[key + ".emd.weight"] = (2 * c, cd); S[key + ".emd.bias"] = (2 * c,)
        else:
            S[key + ".weight"] = (c,); S[key + ".bias"] = (c,)

    def smlp(key, cin, outs, nd, cond=True):
        for j, oc in enumerate(outs):
            S[f"{key}.layers.{3 * j}.weight"] = (oc, cin) + (1,) * nd
            S[f"{key}.layers.{3 * j}.bias"] = (oc,)
            norm(f"{key}.layers.{3 * j + 1}", oc, cond)
            cin = oc

    def pvc(key, cin, cout):
        S[key + ".voxel_layers.0.weight"] = (cout, cin, 3, 3, 3); S[key + ".voxel_layers.0.bias"] = (cout,)
        norm(key + ".voxel_layers.1", cout)
        S[key + ".voxel_layers.4.weight"] = (cout, cout, 3, 3, 3); S[key + ".voxel_layers.4.bias"] = (cout,)
        norm(key + ".voxel_layers.5", cout)
        if use_se:
            S[key + ".voxel_layers.6.fc.0.weight"] = (cout // 8, cout); S[key + ".voxel_layers.6.fc.2.weight"] = (cout, cout // 8)
        smlp(key + ".point_features", cin, [cout], 1)

    for k in ("embedf.0", "embedf.2"):
        S[k + ".weight"] = (E, E); S[k + ".bias"] = (E,)
    if cd > 0:
        for k, (ci, co) in {"mlp1.shared_mlp_0": (ind, cd // 8), "mlp1.shared_mlp_1": (cd // 8, cd // 4),
                            "mlp2.shared_mlp_0": (cd // 2, cd // 2), "mlp2.shared_mlp_1": (cd // 2, cd)}.items():
            S[f"global_pnet.{k}.mlp.0.weight"] = (co, ci, 1, 1); S[f"global_pnet.{k}.mlp.0.bias"] = (co,)
            S[f"global_pnet.{k}.mlp.1.group_norm.weight"] = (co,); S[f"global_pnet.{k}.mlp.1.group_norm.bias"] = (co,)
    if fe != extra:
        ci = ind if extra == 0 else extra
        S["embed_feats.0.weight"] = (fe, ci, 1); S["embed_feats.0.bias"] = (fe,)
        S["embed_feats.1.weight"] = (fe,); S["embed_feats.1.bias"] = (fe,)
        S["embed_feats.3.weight"] = (fe, fe, 1); S["embed_feats.3.bias"] = (fe,)
    cin = fe + ind
    sa_in = []
    L = len(plan["sa"])
    for i, lv in enumerate(plan["sa"]):
        sa_in.append(cin)
        c_feat = cin
        for k in range(lv["n_pvconv"]):
            pvc(f"sa_layers.{i}.{k}", c_feat + (E if (i > 0 and k == 0) else 0), ch[i])
            c_feat = ch[i]
        outs = [ch[i], ch[i + 1]] if i != L - 1 else [ch[i], ch[i], ch[i + 1]]
        key = f"sa_layers.{i}.{lv['n_pvconv']}" if lv["n_pvconv"] > 0 else f"sa_layers.{i}"
        smlp(key + ".mlps.0", c_feat + (E if lv["n_pvconv"] == 0 else 0) + 3, outs, 2)
        cin = outs[-1]
    if str(pvd.get("attention_type") or "linear").lower() == "linear":
        h = plan["heads"]
        S["global_att.to_qkv.weight"] = (3 * h * 32, cin, 1, 1)
        S["global_att.to_out.weight"] = (cin, h * 32, 1, 1); S["global_att.to_out.bias"] = (cin,)
    elif str(pvd.get("attention_type")).lower() == "flash":
        h = plan["heads"]
        S["global_att.to_q.weight"] = (h * 32, cin)
        S["global_att.to_kv.weight"] = (2 * h * 32, cin)
        S["global_att.to_out.weight"] = (cin, h * 32)
    sa_in[0] = fe + ind
    fp_cfg = [((ch[3], ch[3]), ch[3]), ((ch[3], ch[3]), ch[3]), ((ch[3], ch[2]), ch[2]), ((ch[2], ch[2], ch[1]), ch[1])]
    for j, lv in enumerate(plan["fp"]):
        outs, cconv = fp_cfg[j]
        key = f"fp_layers.{j}.0" if lv["n_pvconv"] > 0 else f"fp_layers.{j}"
        smlp(key + ".mlp", cin + sa_in[-1 - j] + E, list(outs), 1)
        cin = outs[-1]
        for k in range(lv["n_pvconv"]):
            pvc(f"fp_layers.{j}.{k + 1}", cin, cconv)
            cin = cconv
    om = pvd.get("out_mlp", 128)
    smlp("classifier.0", cin, [om], 1, cond=False)
    S["classifier.2.weight"] = (m.get("out_dim") or 3, om, 1); S["classifier.2.bias"] = (m.get("out_dim") or 3,)
    return S


def make_state_dict(cfg: dict, seed: int = 0, head_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic backbone weights, one independent stream per key (so the order of construction never
    matters).  Every branch contributes: conv/linear weights ~ U(+-1/sqrt(fan_in)), norm gains 1+0.1 N(0,1),
    biases 0.05 N(0,1), AdaGN ``emd`` biases keep the reference's (1, 0) centre (modules.py:338-339)."""
    import zlib

    sd = {}
    for key, shp in param_shapes(cfg).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        if key.endswith(".weight") and len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            if ".emd." in key:
                bound *= 0.5
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif key.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.05 * torch.randn(shp, generator=g)
            if key.endswith(".emd.bias"):
                t[: shp[0] // 2] += 1.0
        if key.startswith("classifier.2."):
            # head_scale < 1 damps the predicted noise (a converged denoiser on nearly clean input): with purely random
            # weights the T-step map is chaotic (every voxel/FPS/ball-query flip reaches all points through the global
            # conditioning), which makes free-running loop comparisons meaningless beyond a few steps
            t = t * head_scale
        sd[key] = t.float()
    return sd
