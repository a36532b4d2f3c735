"""Generate ``tests/golden/model_*.npz`` by running the REFERENCE's real Python model code in this container.

TEST INFRASTRUCTURE ONLY.  Runs where ``/root/reference`` exists (the build container); the fixtures it writes
are committed, so nothing at test/bench time needs the reference.

What runs: the reference's unmodified ``models/unet_pvc.py`` / ``models/pvcnn.py`` / ``models/modules.py`` /
``models/p2pb.py`` and its six op wrapper files ``third_party/openpoints/models/layers/*.py`` (imported with the
stub recipe of SURVEY.md App. C), with the wrappers' ``pointnet2_cuda`` bound to the CPU restatement
``oracle/ops.py`` (the reference's kernels are CUDA-only; those are pinned separately on the GPU by
``oracle/gen_golden_gpu.py``).  So these fixtures pin everything ABOVE the op layer -- network wiring, layer
semantics, schedule, sampling loop -- of ``oracle/model.py`` against the reference itself, and they check the
parameter inventory (``param_shapes``) key-by-key against the reference's real constructor.

Usage:  python -m oracle.gen_golden            (from the repo root)
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("P2PB_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, ROOT)
from oracle import model as OM  # noqa: E402
from oracle import ops as OO  # noqa: E402


from oracle.ref_import import AttrDict, import_reference as _import_reference, load_ref_cfg  # noqa: E402


def import_reference():
    """SURVEY.md App. C recipe (oracle/ref_import.py), with the op module swapped for the CPU restatement."""
    return _import_reference(OO, REF)


def load_cfg(name, **over):
    cfg = load_ref_cfg(name, REF, **over)
    cfg["gpu"] = "cpu"
    return cfg


def make_input(kind, B, N, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "test_xyz":
        pts = torch.from_numpy(np.loadtxt(f"{REF}/test.xyz").astype(np.float32))[:N]
        pts = pts - pts.mean(0, keepdim=True)
        pts = pts / pts.norm(dim=1).max()
        return pts.t().unsqueeze(0).contiguous()
    x = torch.randn(B, 3, N, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * (1 + 0.05 * torch.randn(B, 1, N, generator=g))
    x = x - x.mean(2, keepdim=True)
    return (x / x.norm(dim=1).amax(dim=1)[:, None, None]).contiguous()


CASES = [
    # name, config, overrides, input kind, B, N, T, x_cond channels, head_scale
    ("pvds_cfg1", "PVDS_PUNet", {}, "test_xyz", 1, 1024, 5, 0, 1.0),
    ("pvds_b2", "PVDS_PUNet", {}, "synth", 2, 2048, 2, 0, 1.0),
    ("pvdl_xyz", "PVDL_SNPP", {"data.npoints": 512, "model.extra_feature_channels": 0}, "synth", 1, 512, 2, 0, 1.0),
    ("pvdl_rgb", "PVDL_SNPP", {"data.npoints": 512, "model.extra_feature_channels": 3}, "synth", 1, 512, 1, 3, 1.0),
    ("pvdl_dino", "PVDL_SNPP", {"data.npoints": 512, "model.extra_feature_channels": 387}, "synth", 1, 512, 1, 387, 1.0),
    # damped noise head: well-conditioned free-running loops (see oracle/model.py make_state_dict)
    ("pvds_cfg1_damped", "PVDS_PUNet", {}, "test_xyz", 1, 1024, 5, 0, 0.02),
    ("pvds_t30_damped", "PVDS_PUNet", {}, "synth", 2, 2048, 30, 0, 0.02),
    # un-damped T=30 chain: the teacher-forced test advances the reference's state one step at a time, so chaos does not
    # matter and the per-step displacement is large enough for a RELATIVE bound (an identity step fails)
    ("pvds_t30", "PVDS_PUNet", {}, "synth", 2, 2048, 30, 0, 1.0),
    # attention_type: flash -> modules.Attention / Attend at the bottleneck (unet_pvc.py:98-99, 238-241)
    ("pvds_flash", "PVDS_PUNet", {"model.PVD.attention_type": "flash"}, "synth", 2, 1024, 2, 0, 1.0),
    ("pvdl_flash", "PVDL_SNPP", {"data.npoints": 512, "model.extra_feature_channels": 0, "model.PVD.attention_type": "flash"},
     "synth", 1, 512, 1, 0, 1.0),
]


def main():
    only = set(sys.argv[1:])
    os.makedirs(OUT, exist_ok=True)
    PVCNN2Unet, P2PB = import_reference()
    torch.set_grad_enabled(False)
    report = []
    for name, cfgname, over, kind, B, N, T, F, hs in CASES:
        if only and name not in only:
            continue
        cfg = load_cfg(cfgname, **over)
        acfg = AttrDict.wrap(copy.deepcopy(cfg))
        acfg.model.ema = False
        net = PVCNN2Unet(acfg)
        ref_shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        ora_shapes = OM.param_shapes(cfg)
        assert ref_shapes == ora_shapes, (
            name, set(ref_shapes) ^ set(ora_shapes), [k for k in ref_shapes if ora_shapes.get(k) != ref_shapes[k]][:5])
        sd = OM.make_state_dict(cfg, seed=0, head_scale=hs)
        net.load_state_dict(sd, strict=True)
        model = P2PB(acfg, net)
        x = make_input(kind, B, N, seed=1)
        xc = None
        if F:
            g = torch.Generator().manual_seed(2)
            xc = torch.cat([torch.rand(B, 3, N, generator=g), torch.randn(B, F - 3, N, generator=g)], 1) if F > 3 \
                else torch.rand(B, 3, N, generator=g)
            xc = xc.half().float()  # stored as fp16 in the fixture: make the values exactly representable
        # ---- reference: one forward at the first sampling step, then the T-step loop
        sched = OM.build_schedule(cfg)
        for k in ("std_fwd", "std_bwd", "std_sb", "mu_x0", "mu_x1", "betas", "noise_levels"):
            assert torch.equal(getattr(model, k).cpu(), sched[k]), k
        step_hi = OM.space_indices(cfg["diffusion"]["timesteps"], T + 1)[-1]
        nl = model.noise_levels[torch.full((B,), step_hi, dtype=torch.long)]
        net.eval()
        eps_ref = net(x, nl, x_cond=xc)
        out_ref = model.sample(x_start=x, x_cond=xc, steps=T, log_count=T, verbose=False, use_ema=False)
        # ---- oracle restatement on the same weights / inputs
        eps_ora = OM.unet_forward(sd, cfg, x, nl, xc)
        out_ora = OM.sample(sd, cfg, x, xc, steps=T, log_count=T)
        d_eps = (eps_ref - eps_ora).abs().max().item()
        d_x = (out_ref["x_pred"] - out_ora["x_pred"]).abs().max().item()
        d_chain = (out_ref["x_chain"] - out_ora["x_chain"]).abs().max().item()
        report.append((name, d_eps, d_x, d_chain, eps_ref.abs().max().item()))
        print(f"{name}: |eps_ref-eps_oracle|max={d_eps:.3e}  |x_pred diff|max={d_x:.3e}  chain={d_chain:.3e} "
              f"(|eps|max={eps_ref.abs().max():.3f})")
        np.savez_compressed(
            os.path.join(OUT, f"model_{name}.npz"),
            x_start=x.numpy(), x_cond=(xc.numpy().astype(np.float16) if xc is not None else np.zeros(0)),
            eps=eps_ref.numpy(), x_pred=out_ref["x_pred"].numpy(), x_chain=out_ref["x_chain"].numpy(),
            noise_level=nl.numpy(), T=T, cfg_name=cfgname, overrides=yaml.safe_dump(over),
            n_params=len(ref_shapes), head_scale=hs)
    # schedule fixtures for both beta_end settings (p2pb.py:93-130) and the step grids (p2pb.py:16-40)
    for cfgname in ("PVDS_PUNet", "PVDL_SNPP"):
        if only:
            break
        cfg = load_cfg(cfgname)
        s = OM.build_schedule(cfg)
        np.savez_compressed(os.path.join(OUT, f"schedule_{cfgname}.npz"), **{k: v.numpy() for k, v in s.items()},
                            steps5=np.array(OM.space_indices(1000, 6)), steps30=np.array(OM.space_indices(1000, 31)))
    return report


if __name__ == "__main__":
    main()
