"""R-GPU: run the UNMODIFIED reference (its ``models/*.py`` over its own compiled ``pointnet2_batch_cuda``) on a B200.

TEST / BASELINE INFRASTRUCTURE ONLY.  Needs ``baseline/_ref`` (``oracle/snapshot_ref.py``) and ``oracle/_ref/*.so``
(``oracle/build_ref.py``); both travel to the GPU box with the gpurun snapshot.  Writes ``gpurun_out/rgpu_*.npz`` (copy to
``tests/golden/``) and ``gpurun_out/rgpu_report.json`` (copy to ``profiles/``):

  * the reference's own throughput on BASELINE config 2 (PVDS, 64 x 2048, T=30) and config 3 (PVDL, N=8192), PyTorch
    defaults (cuDNN convolutions in TF32, matmul fp32) -- the apples-to-apples GPU baseline of BASELINE.md §2;
  * its NOISE FLOOR: Chamfer(run 1, run 2) with identical inputs (fp32 atomics in ``avg_voxelize_kernel``) and
    Chamfer(TF32 convs, fp32 convs), for the un-damped and the damped-head seeded checkpoints -- the yard-stick every T-step
    tolerance in ``tests/`` is judged against;
  * goldens of the real reference on the real GPU: one evaluation at the bench batch (64 patches), the damped T=30 result at
    the bench batch, a teacher-forcing chain for PVDL at N=8192, and the reference's own single-step GPU-vs-CPU spread along
    the committed CPU chain (``tests/golden/model_pvds_t30.npz``).

Usage (GPU box):  python -m oracle.gen_golden_rgpu [--quick]
"""
from __future__ import annotations

import copy
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import model as OM  # noqa: E402
from oracle.ref_import import AttrDict, import_reference, load_ref_cfg, load_ref_extension, reference_root  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
DRY = os.environ.get("P2PB_RGPU_DRYRUN") == "1"     # CPU plumbing check of this script (oracle ops, tiny sizes); never a golden
DEV = "cpu" if DRY else "cuda:0"


def chamfer(cham, a, b):
    """metrics/metrics.py:56-83 (calculate_cd_cuda) on the reference's own chamfer_3D extension: [B,3,N] x2 -> [B] tensor."""
    if DRY:
        from oracle import ops as OO

        return torch.tensor(OO.calculate_cd(a.cpu(), b.cpu()))
    p1 = a.transpose(1, 2).contiguous()
    p2 = b.transpose(1, 2).contiguous()
    B, n, _ = p1.shape
    m = p2.shape[1]
    d1 = torch.zeros(B, n, device=a.device)
    d2 = torch.zeros(B, m, device=a.device)
    i1 = torch.zeros(B, n, dtype=torch.int32, device=a.device)
    i2 = torch.zeros(B, m, dtype=torch.int32, device=a.device)
    cham.forward(p1, p2, d1, d2, i1, i2)
    return (d1.mean(1) + d2.mean(1)).cpu()


def set_tf32(on: bool):
    """PyTorch defaults (what the reference runs with): cuDNN TF32 on, matmul TF32 off.  off = everything fp32."""
    torch.backends.cudnn.allow_tf32 = bool(on)
    torch.backends.cuda.matmul.allow_tf32 = False


def build_model(PVCNN2Unet, P2PB, cfg_dict, head_scale):
    acfg = AttrDict.wrap(copy.deepcopy(cfg_dict))
    acfg.gpu = DEV
    acfg.model.ema = False
    net = PVCNN2Unet(acfg)
    sd = OM.make_state_dict(cfg_dict, seed=0, head_scale=head_scale)
    net.load_state_dict(sd, strict=True)
    model = P2PB(acfg, net)
    model.eval()            # the reference's eval() returns None (train_utils.py)
    return model


def timed_sample(model, x, xc, T, reps):
    if DRY:
        t0 = time.time()
        for _ in range(reps):
            out = model.sample(x_start=x, x_cond=xc, steps=T, log_count=1, verbose=False, use_ema=False)
        return (time.time() - t0) * 1e3 / reps, out
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = model.sample(x_start=x, x_cond=xc, steps=T, log_count=1, verbose=False, use_ema=False)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def stats(t):
    t = t.double()
    return {"mean": float(t.mean()), "max": float(t.max()), "median": float(t.median())}


def main():
    quick = "--quick" in sys.argv
    os.makedirs(OUT, exist_ok=True)
    import bench
    from tests.helpers import patch_input

    ref = reference_root()
    if DRY:
        from oracle import ops as ext
        cham = None
    else:
        ext = load_ref_extension("pointnet2_batch_cuda")
        cham = load_ref_extension("chamfer_3D")
    PVCNN2Unet, P2PB = import_reference(ext, ref)
    torch.set_grad_enabled(False)
    report = {"gpu": "dry-run (cpu)" if DRY else torch.cuda.get_device_name(0), "torch": torch.__version__, "reference_root": ref,
              "tf32": "PyTorch defaults: cudnn.allow_tf32=True, cuda.matmul.allow_tf32=False"}
    gold = {}

    # ------------------------------------------------------------------ config 2: PVDS, 64 x 2048, T=30 (the bench input)
    cfg = load_ref_cfg("PVDS_PUNet", ref)
    B, N, T = (8 if quick else 64), 2048, 30
    if DRY:
        B, T = 1, 2
    x = bench.synth_patches(64, N, seed=1000)[:B].to(DEV)
    for tag, hs in (("undamped", 1.0), ("damped", 0.02)):
        model = build_model(PVCNN2Unet, P2PB, cfg, hs)
        set_tf32(True)
        model.sample(x_start=x, steps=2, log_count=1, verbose=False, use_ema=False)          # warm-up (cuDNN autotune etc.)
        ms, o1 = timed_sample(model, x, None, T, 1)
        ms2, o2 = timed_sample(model, x, None, T, 2)
        set_tf32(False)
        _, o3 = timed_sample(model, x, None, T, 1)
        _, o4 = timed_sample(model, x, None, T, 1)
        a, b, c, d = o1["x_pred"], o2["x_pred"], o3["x_pred"], o4["x_pred"]
        moved = (c - x).abs().mean().item()
        r = {
            "B": B, "N": N, "T": T, "head_scale": hs, "ms_per_sample_call": ms2, "patches_per_s": B / (ms2 / 1e3),
            "cd_run1_vs_run2_tf32": stats(chamfer(cham, a, b)), "cd_run1_vs_run2_fp32": stats(chamfer(cham, c, d)),
            "cd_tf32_vs_fp32": stats(chamfer(cham, a, c)), "cd_do_nothing": stats(chamfer(cham, x, c)),
            "mean_abs_diff_run1_vs_run2_tf32": (a - b).abs().mean().item(), "mean_abs_diff_tf32_vs_fp32": (a - c).abs().mean().item(),
            "mean_abs_moved": moved,
        }
        report[f"cfg2_{tag}"] = r
        print(tag, json.dumps(r), flush=True)
        if tag == "damped":
            gold["cfg2_damped_x_pred_fp32"] = c.cpu().numpy()
            gold["cfg2_damped_x_pred_tf32"] = a.cpu().numpy()
        else:
            # one evaluation at the bench batch (first sampling step), fp32 -- continuous compare target for the engine at B=64
            step_hi = OM.space_indices(1000, T + 1)[-1]
            nl = model.noise_levels[torch.full((B,), step_hi, dtype=torch.long, device=DEV)]
            model.model.eval()      # ddpm_sampling leaves the net in train mode (p2pb.py:333)
            eps32 = model.model(x, nl)
            set_tf32(True)
            eps_tf = model.model(x, nl)
            eps_tf2 = model.model(x, nl)
            gold["cfg2_eps_fp32"] = eps32.cpu().numpy()
            gold["cfg2_noise_level"] = nl.cpu().numpy()
            report["cfg2_eps_tf32_vs_fp32"] = {"mean": (eps_tf - eps32).abs().mean().item(), "max": (eps_tf - eps32).abs().max().item(),
                                               "run_to_run_max": (eps_tf - eps_tf2).abs().max().item()}
            print("eps tf32 vs fp32", report["cfg2_eps_tf32_vs_fp32"], flush=True)
        del model
        torch.cuda.empty_cache()
    gold["cfg2_B"] = B

    # ------------------------------------------------------------------ the reference's own single-step GPU-vs-CPU spread
    z = np.load(os.path.join(ROOT, "tests", "golden", "model_pvds_t30.npz"))
    Tc = int(z["T"])
    n_tf = 2 if DRY else Tc
    chain = torch.from_numpy(z["x_chain"]).to(DEV)
    x0 = torch.from_numpy(z["x_start"]).to(DEV)
    model = build_model(PVCNN2Unet, P2PB, cfg, 1.0)
    rev = OM.space_indices(1000, Tc + 1)[::-1]
    rel = {"tf32": [], "fp32": []}
    model.model.eval()
    for mode in ("tf32", "fp32"):
        set_tf32(mode == "tf32")
        for s, (prev, step) in enumerate(list(zip(rev[1:], rev[:-1]))[:n_tf]):
            before = x0 if s == 0 else chain[:, Tc - s]
            after = chain[:, Tc - 1 - s]
            st = torch.full((before.shape[0],), step, device=DEV, dtype=torch.long)
            eps = model.model(before, model.noise_levels[st])
            px0 = model.compute_pred_x0_from_eps(st, before, eps, False)
            got = model.p_posterior(prev, step, before, px0)
            rel[mode].append(((got - after).abs().mean() / (after - before).abs().mean()).item())
    report["teacher_forced_rgpu_vs_rcpu_rel_err"] = {k: {"max": max(v), "mean": float(np.mean(v))} for k, v in rel.items()}
    print("teacher-forced R-GPU vs R-CPU relative error", report["teacher_forced_rgpu_vs_rcpu_rel_err"], flush=True)
    del model
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ config 3: PVDL, N=8192, xyz
    cfgL = load_ref_cfg("PVDL_SNPP", ref, **{"data.npoints": 1024 if DRY else 8192, "model.extra_feature_channels": 0})
    model = build_model(PVCNN2Unet, P2PB, cfgL, 1.0)
    set_tf32(False)
    xl = patch_input(2, 1024 if DRY else 8192, seed=5).to(DEV)
    out = model.sample(x_start=xl, steps=5, log_count=5, verbose=False, use_ema=False)
    gold["pvdl8192_x_start"] = xl.cpu().numpy()
    gold["pvdl8192_x_chain"] = out["x_chain"].cpu().numpy()
    step_hi = OM.space_indices(1000, 6)[-1]
    nl = model.noise_levels[torch.full((2,), step_hi, dtype=torch.long, device=DEV)]
    model.model.eval()
    gold["pvdl8192_eps_fp32"] = model.model(xl, nl).cpu().numpy()
    gold["pvdl8192_noise_level"] = nl.cpu().numpy()
    set_tf32(True)
    Bl = 1 if DRY else (4 if quick else 32)
    xb = patch_input(Bl, 1024 if DRY else 8192, seed=7).to(DEV)
    model.sample(x_start=xb, steps=2, log_count=1, verbose=False, use_ema=False)
    ms, _ = timed_sample(model, xb, None, 2 if DRY else 30, 1)
    report["cfg3_pvdl_xyz"] = {"B": Bl, "N": 8192, "T": 30, "ms_per_sample_call": ms, "patches_per_s": Bl / (ms / 1e3)}
    print("cfg3", report["cfg3_pvdl_xyz"], flush=True)
    del model
    torch.cuda.empty_cache()

    # the reference's own noise floor travels with the goldens so that the tests state their tolerances against it
    d = report["cfg2_damped"]
    gold["floor_cd_tf32_vs_fp32_mean"], gold["floor_cd_tf32_vs_fp32_max"] = d["cd_tf32_vs_fp32"]["mean"], d["cd_tf32_vs_fp32"]["max"]
    gold["floor_cd_run_to_run_mean"], gold["floor_cd_run_to_run_max"] = d["cd_run1_vs_run2_tf32"]["mean"], d["cd_run1_vs_run2_tf32"]["max"]
    gold["floor_teacher_forced_rel_tf32_max"] = report["teacher_forced_rgpu_vs_rcpu_rel_err"]["tf32"]["max"]
    name = "rgpu_dryrun" if DRY else "rgpu"
    np.savez_compressed(os.path.join(OUT, f"{name}_golden.npz"), **gold)
    json.dump(report, open(os.path.join(OUT, f"{name}_report.json"), "w"), indent=1)
    print("wrote", os.path.join(OUT, f"{name}_golden.npz"), os.path.getsize(os.path.join(OUT, f"{name}_golden.npz")))


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"done in {time.time() - t0:.0f} s")
