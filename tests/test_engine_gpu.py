"""GPU parity of the fused engine (hand-written kernels only, CUDA graph) against
  * the golden outputs of the reference's REAL model code (tests/golden/model_*.npz, fp32 CPU) and
  * the eager reference-shaped path on the same device.
Tolerance: the engine's contractions are TF32 tensor-core MMAs with fp32 accumulation -- the arithmetic the reference's
own convolutions use on a GPU (cuDNN TF32 default) -- so one network evaluation is compared to the fp32 golden with
mean |err| <= 2e-3 and max |err| <= 3e-2 on eps of O(1); the T-step loop is compared set-wise (Chamfer, the
north-star bound 1e-5) because discrete ops (voxel rounding, FPS, ball query) flip on rounding-level differences."""
import os

import numpy as np
import pytest
import torch
import yaml

from tests.helpers import load_cfg, patch_input
from tests.test_model_gpu import build

pytestmark = pytest.mark.gpu


def _golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    cfg = load_cfg(str(z["cfg_name"]), **(yaml.safe_load(str(z["overrides"])) or {}))
    return z, cfg


@pytest.mark.parametrize("name", ["pvds_cfg1", "pvds_b2", "pvdl_xyz", "pvdl_rgb", "pvdl_dino", "pvds_flash", "pvdl_flash"])
def test_engine_single_evaluation_vs_reference_golden(golden_dir, name):
    from p2pb_b200.engine import get_engine

    z, cfg = _golden(golden_dir, name)
    model, _ = build(cfg, backend="engine")
    x = torch.from_numpy(z["x_start"]).cuda()
    xc = torch.from_numpy(z["x_cond"].astype(np.float32)).cuda() if z["x_cond"].size else None
    eng = get_engine(model, model.model, x.shape, None if xc is None else xc.shape)
    B, _, N = x.shape
    nl = float(z["noise_level"][0])
    with torch.no_grad():
        if xc is not None:
            eng.prepare_cond(xc)
        sin = eng.time_embedding(nl, None)[None].expand(B, -1).contiguous()
        eps_rows = eng.evaluate(x.contiguous(), sin)
    torch.cuda.synchronize()
    eps = eps_rows[:, :3].reshape(B, N, 3).permute(0, 2, 1).cpu().numpy()
    err = np.abs(eps - z["eps"])
    print(f"{name}: engine vs reference golden eps: mean|err|={err.mean():.3e} max|err|={err.max():.3e} |eps|max={np.abs(z['eps']).max():.2f}")
    assert err.mean() <= 2e-3 and err.max() <= 3e-2, (err.mean(), err.max())


def _teacher_forced(eng, x_start, chain, T, rel_bound, tag):
    """Advance the REFERENCE's state x_t by one engine step and compare with the reference's x_{t-1}, for every step.
    The bound is RELATIVE to that step's displacement (mean|err| <= rel_bound * mean|x_{t-1} - x_t|): a step that returned its
    input has relative error 1 and fails, whatever the size of the step."""
    from p2pb_b200.p2pb import space_indices

    rev = space_indices(1000, T + 1)[::-1]
    worst = 0.0
    for s, (prev, step) in enumerate(zip(rev[1:], rev[:-1])):
        before = x_start if s == 0 else chain[:, T - s]
        after_ref = chain[:, T - 1 - s]
        xs, _ = eng.sample(before.contiguous(), None, [(prev, step)], [prev], False)
        err = (xs[:, 0] - after_ref).abs().mean().item()
        disp = (after_ref - before).abs().mean().item()
        worst = max(worst, err / disp)
        assert err <= rel_bound * disp, (tag, s, err, disp, err / disp)
    print(f"{tag} teacher-forced: worst per-step mean|err| / mean|dx| over {T} steps = {worst:.3e} (bound {rel_bound})")
    return worst


# Measured on B200 (profiles/r02_rgpu_report.json): the reference's OWN GPU path (cuDNN TF32 convolutions, its compiled
# kernels) differs from its fp32 CPU run by 0.18 % (mean over steps) / 0.25 % (worst step) of the step displacement; its fp32
# GPU run by 3e-6.  The engine (10-bit-mantissa operands everywhere a contraction runs, not only in the convolutions) is held
# to 1 %: four times the reference's own TF32 spread, two orders of magnitude below "did nothing".
TEACHER_FORCED_REL = 0.01


@pytest.mark.parametrize("name", ["pvds_cfg1", "pvds_t30", "pvds_t30_damped"])
def test_engine_teacher_forced_steps_vs_reference_golden(golden_dir, name):
    """Every step of the loop (T=5 on config 1 = first 1024 points of the reference's test.xyz; T=30 on 2 x 2048 points with
    the UN-damped and the damped seeded checkpoint), teacher-forced against the chain of the reference's real model code."""
    from p2pb_b200.engine import get_engine

    z, cfg = _golden(golden_dir, name)
    model, _ = build(cfg, backend="engine", head_scale=float(z["head_scale"]))
    T = int(z["T"])
    chain = torch.from_numpy(z["x_chain"]).cuda()          # [B, T, 3, N], index 0 = final state
    x = torch.from_numpy(z["x_start"]).cuda()
    eng = get_engine(model, model.model, x.shape, None)
    _teacher_forced(eng, x, chain, T, TEACHER_FORCED_REL, name)


def test_engine_teacher_forced_pvdl_8192_vs_rgpu(golden_dir):
    """PVDL at N = 8192 (BASELINE configs 3-4), T = 5 loop, against the chain of the UNMODIFIED reference run on a B200
    (fp32; oracle/gen_golden_rgpu.py), plus one evaluation against its eps."""
    from p2pb_b200.engine import get_engine

    z = np.load(os.path.join(golden_dir, "rgpu_golden.npz"))
    cfg = load_cfg("PVDL_SNPP", **{"data.npoints": 8192, "model.extra_feature_channels": 0})
    model, _ = build(cfg, backend="engine")
    x = torch.from_numpy(z["pvdl8192_x_start"]).cuda()
    chain = torch.from_numpy(z["pvdl8192_x_chain"]).cuda()
    eng = get_engine(model, model.model, x.shape, None)
    _teacher_forced(eng, x, chain, 5, TEACHER_FORCED_REL, "pvdl8192")
    eps = _one_evaluation(eng, x, float(z["pvdl8192_noise_level"][0]))
    err = (eps - torch.from_numpy(z["pvdl8192_eps_fp32"]).cuda()).abs()
    print(f"PVDL N=8192 engine vs R-GPU fp32 eps: mean|err|={err.mean():.3e} max|err|={err.max():.3e}")
    assert err.mean().item() <= 2e-3 and err.max().item() <= 5e-2


def _one_evaluation(eng, x, noise_level):
    B, _, N = x.shape
    with torch.no_grad():
        sin = eng.time_embedding(noise_level, None)[None].expand(B, -1).contiguous()
        eps_rows = eng.evaluate(x.contiguous(), sin)
    torch.cuda.synchronize()
    return eps_rows[:, :3].reshape(B, N, 3).permute(0, 2, 1)


def test_engine_bench_config_vs_rgpu(golden_dir):
    """The BENCHMARKED configuration itself (bench.py: 64 synthetic 2048-point patches, PVDS, T = 30) against the UNMODIFIED
    reference run on a B200 (oracle/gen_golden_rgpu.py -> tests/golden/rgpu_golden.npz).  At 64 patches the kernels take
    the CTA-pair / balanced-tile-range code paths the small goldens do not reach.
      (1) one evaluation, un-damped weights, continuous compare with the reference's fp32 eps (same tolerance as the small
          goldens; the reference's own TF32 path differs from it by 2.9e-4 mean / 2.3e-3 max, 1.5e-3 max run to run);
      (2) free-running T = 30 on the damped-head checkpoint: Chamfer(engine, R-GPU fp32) per patch within
          max(1e-5, 2 x the reference's own Chamfer(TF32, fp32)) -- mean and worst patch.  The reference's floor, measured:
          mean 1.9e-6 / worst patch 7.3e-6 (TF32 vs fp32), 1.1e-6 / 6.6e-6 (run to run, fp32 atomics).
    Un-damped free-running T = 30 is not comparable for ANY implementation: the reference's own two runs differ by Chamfer
    1.2e-2 (mean over patches; do-nothing 7.5e-2), profiles/r02_rgpu_report.json."""
    import bench
    from p2pb_b200 import ops
    from p2pb_b200.engine import get_engine

    z = np.load(os.path.join(golden_dir, "rgpu_golden.npz"))
    B = int(z["cfg2_B"])
    cfg = load_cfg("PVDS_PUNet")
    x = bench.synth_patches(64, 2048, seed=1000)[:B].cuda()
    model, _ = build(cfg, backend="engine")
    eng = get_engine(model, model.model, x.shape, None)
    eps = _one_evaluation(eng, x, float(z["cfg2_noise_level"][0]))
    err = (eps - torch.from_numpy(z["cfg2_eps_fp32"]).cuda()).abs()
    print(f"B={B} engine vs R-GPU fp32 eps: mean|err|={err.mean():.3e} max|err|={err.max():.3e}")
    assert err.mean().item() <= 2e-3 and err.max().item() <= 3e-2
    model, _ = build(cfg, backend="engine", head_scale=0.02)
    out = model.sample(x_start=x, steps=30, log_count=1, verbose=False)["x_pred"]
    ref = torch.from_numpy(z["cfg2_damped_x_pred_fp32"]).cuda()
    cd = torch.tensor(ops.calculate_cd(out, ref))
    cd0 = torch.tensor(ops.calculate_cd(x, ref))
    b_mean = max(1e-5, 2 * float(z["floor_cd_tf32_vs_fp32_mean"]))
    b_max = max(1e-5, 2 * float(z["floor_cd_tf32_vs_fp32_max"]))
    print(f"B={B} T=30 damped: chamfer(engine, R-GPU fp32) mean={cd.mean():.3e} max={cd.max():.3e} (bounds {b_mean:.2e} / {b_max:.2e}; "
          f"reference TF32-vs-fp32 floor {float(z['floor_cd_tf32_vs_fp32_mean']):.2e} / {float(z['floor_cd_tf32_vs_fp32_max']):.2e}; "
          f"do-nothing {cd0.mean():.3e})")
    assert cd.mean().item() <= b_mean and cd.max().item() <= b_max, (cd.mean().item(), cd.max().item())


@pytest.mark.parametrize("name", ["pvds_cfg1_damped", "pvds_t30_damped"])
def test_engine_free_running_loop_vs_reference_golden(golden_dir, name):
    """Free-running T-step loop through P2PB.sample (engine, one CUDA graph) on the damped-head checkpoint vs the
    reference's real P2PB.sample (fp32 CPU).  Bound: max(1e-5, 2 x the reference's own TF32-vs-fp32 Chamfer on a B200 for a
    T = 30 loop of this checkpoint, worst patch 7.3e-6 -- tests/golden/rgpu_golden.npz) = 1.47e-5; T = 5: the north-star 1e-5."""
    from p2pb_b200 import ops

    z, cfg = _golden(golden_dir, name)
    fl = np.load(os.path.join(golden_dir, "rgpu_golden.npz"))
    model, _ = build(cfg, backend="engine", head_scale=float(z["head_scale"]))
    x = torch.from_numpy(z["x_start"]).cuda()
    T = int(z["T"])
    out = model.sample(x_start=x, steps=T, log_count=T, verbose=False)
    out2 = model.sample(x_start=x, steps=T, log_count=T, verbose=False)      # graph replay: bit-identical
    assert torch.equal(out["x_pred"], out2["x_pred"])
    assert out["x_chain"].shape == z["x_chain"].shape
    ref = torch.from_numpy(z["x_pred"]).cuda()
    cd = ops.calculate_cd(out["x_pred"], ref)
    diff = (out["x_pred"] - ref).abs()
    moved = (ref - x).abs().mean().item()
    bound = 1e-5 if T <= 5 else max(1e-5, 2 * float(fl["floor_cd_tf32_vs_fp32_max"]))
    print(f"{name}: T={T} chamfer={max(cd):.3e} (bound {bound:.2e}) mean|diff|={diff.mean():.3e} max|diff|={diff.max():.3e} "
          f"(mean |x_pred-x_start|={moved:.3e})")
    assert max(cd) <= bound, cd
    assert diff.mean().item() < 0.1 * moved + 1e-4


def test_engine_vs_eager_batch(golden_dir):
    """B=4 patches, N=2048, T=3: engine vs the eager path (our ops + torch fp32 library layers) on the same GPU.
    Sanity check of the batched path (both sides differ from each other by TF32 rounding, amplified over the steps by
    the discrete ops; the parity statements are the golden-vector tests above)."""
    from p2pb_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    cfg = load_cfg("PVDS_PUNet")
    model, _ = build(cfg, backend="engine", head_scale=0.02)
    x = patch_input(4, 2048, seed=11).cuda()
    a = model.sample(x_start=x, steps=3, log_count=3, verbose=False, backend="engine")
    b = model.sample(x_start=x, steps=3, log_count=3, verbose=False, backend="eager")
    cd = ops.calculate_cd(a["x_pred"], b["x_pred"])
    diff = (a["x_pred"] - b["x_pred"]).abs()
    print(f"engine vs eager: chamfer={max(cd):.3e} mean|diff|={diff.mean():.3e} max|diff|={diff.max():.3e}")
    torch.backends.cudnn.allow_tf32 = True
    assert max(cd) < 1e-4 and diff.mean().item() < 2e-3
    assert a["x_chain"].shape == b["x_chain"].shape == (4, 3, 3, 2048)


def test_denoise_object_entry_point_end_to_end(tmp_path):
    """denoise_object.py CLI on a synthetic 6k-point cloud with a seeded checkpoint in the reference's format."""
    import yaml as _yaml

    import denoise_object as D
    from p2pb_b200.config import load_yaml
    from p2pb_b200.model_loader import save_checkpoint, seeded_state_dict
    from p2pb_b200.p2pb import P2PB
    from p2pb_b200.unet_pvc import PVCNN2Unet

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = load_yaml(os.path.join(root, "p2pb_b200", "configs", "PVDS_PUNet.yaml"))
    (tmp_path / "opt.yaml").write_text(_yaml.safe_dump(cfg.to_dict()))
    cfg.gpu = "cpu"
    net = PVCNN2Unet(cfg)
    net.load_state_dict(seeded_state_dict(net, 0, head_scale=0.02))
    save_checkpoint(str(tmp_path / "step_0.pth"), P2PB(cfg, net), step=0)
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(6144, 3, generator=g)
    pts = pts / pts.norm(dim=1, keepdim=True) + 0.02 * torch.randn(6144, 3, generator=g)
    np.savetxt(str(tmp_path / "in.xyz"), pts.numpy())
    D.sample(D.parse_args(["--data_path", str(tmp_path / "in.xyz"), "--save_path", str(tmp_path / "out.xyz"),
                           "--model_path", str(tmp_path / "step_0.pth"), "--steps", "3"]))
    out = np.loadtxt(str(tmp_path / "out.xyz"))
    assert out.shape == (6144, 3) and np.isfinite(out).all()
    assert np.abs(np.linalg.norm(out, axis=1) - 1.0).mean() < 0.1      # still the noisy unit sphere, moved a little
    # ---- value check: the same pipeline with the patch extraction (seed FPS, kNN-2048) and the merge FPS done by the CPU restatement
    # of the reference's host code (denoise_object.py:64-122; models/evaluation.py:297-311 FPS from index 0, truncated; pytorch3d
    # knn_points contract: K nearest, ascending) around the SAME network and the same torch normalisation.  The FPS / kNN kernels are
    # bit-exact to the oracle and the engine is deterministic, so the written cloud must equal the oracle-built one to the precision
    # of the "%8f" text format.
    from oracle import ops as OO
    from p2pb_b200.model_loader import load_diffusion

    def o_fps(pcls, num_pnts):
        idx = OO.furthest_point_sampling_forward(pcls.cpu().transpose(1, 2).contiguous().float(), num_pnts).long().to(pcls.device)
        return torch.gather(pcls, 1, idx.unsqueeze(-1).expand(-1, -1, 3)), [i for i in idx]

    def o_knn(seeds, pcl, K):
        idx, _ = OO.knn_points(seeds.cpu(), pcl.cpu(), K)
        return pcl[idx.long().to(pcl.device)]

    a = D.parse_args(["--data_path", str(tmp_path / "in.xyz"), "--save_path", str(tmp_path / "out2.xyz"),
                      "--model_path", str(tmp_path / "step_0.pth"), "--steps", "3"])
    model, _ = load_diffusion(a)
    model.eval()
    pcl = torch.tensor(np.loadtxt(str(tmp_path / "in.xyz")), dtype=torch.float32)
    pcl, center, scale_n = D.normalize_unit_sphere(pcl)
    orig = (D.farthest_point_sampling, D.knn_patches)
    D.farthest_point_sampling, D.knn_patches = o_fps, o_knn
    try:
        den = D.patch_based_denoise(model, pcl.to(a.gpu), patch_size=2048, seed_k=a.k, cfg=a).cpu()
    finally:
        D.farthest_point_sampling, D.knn_patches = orig
    ref = (den * scale_n + center).numpy()
    assert np.abs(out - ref).max() <= 1.5e-6, np.abs(out - ref).max()


@pytest.mark.parametrize("extra", [0, 3])
def test_engine_pvdl_8192_vs_eager(extra):
    """PVDL at the bench shapes of configs 3-4 (N=8192, data.npoints=8192; xyz-only and xyz+RGB), B=2, one network
    evaluation: fused engine vs the eager fp32 path on the same device (same discrete geometry -> continuous compare)."""
    from p2pb_b200.engine import get_engine

    cfg = load_cfg("PVDL_SNPP", **{"data.npoints": 8192, "model.extra_feature_channels": extra})
    model, _ = build(cfg, backend="engine")
    B, N = 2, 8192
    x = patch_input(B, N, seed=5).cuda()
    g = torch.Generator().manual_seed(6)
    xc = torch.rand(B, extra, N, generator=g).cuda() if extra else None
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    nl = model.noise_levels[torch.full((B,), 999, device="cuda", dtype=torch.long)]
    with torch.no_grad():
        ref = model.model(x, nl, x_cond=xc)
        eng = get_engine(model, model.model, x.shape, None if xc is None else xc.shape)
        if xc is not None:
            eng.prepare_cond(xc)
        sin = eng.time_embedding(float(nl[0].item()), None)[None].expand(B, -1).contiguous()
        eps = eng.evaluate(x.contiguous(), sin)[:, :3].reshape(B, N, 3).permute(0, 2, 1)
    torch.backends.cudnn.allow_tf32 = True
    err = (eps - ref).abs()
    print(f"PVDL N=8192 extra={extra}: engine vs eager fp32: mean|err|={err.mean():.3e} max|err|={err.max():.3e} |eps|max={ref.abs().max():.2f}")
    assert err.mean().item() <= 2e-3 and err.max().item() <= 5e-2


def test_voxelize_sparse_equals_dense_and_clears():
    """The occupied-voxels-only voxelisation (engine default) writes bit-identical rows to the dense kernel, leaves
    every other row untouched (zero), and its clear pass restores the all-zero grid."""
    import ctypes

    from p2pb_b200 import dense
    from p2pb_b200._lib import call

    vp = ctypes.c_void_p
    p = lambda t: vp(t.data_ptr()) if t is not None else vp(0)
    s = vp(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(11)
    B, N, r, Cf, E, Cp = 3, 1024, 16, 35, 64, 128
    coords = torch.randn(B, 3, N, device="cuda", generator=g)
    feat = torch.randn(B * N, 64, device="cuda", generator=g)        # ldf = 64 > Cf
    temb = torch.randn(B, E, device="cuda", generator=g)
    nc = torch.empty(B, 3, N, device="cuda")
    ind = torch.empty(B, N, dtype=torch.int32, device="cuda")
    order = torch.empty(B, N, dtype=torch.int32, device="cuda")
    start = torch.empty(B, r ** 3, dtype=torch.int32, device="cuda")
    cnt = torch.empty(B, r ** 3, dtype=torch.int32, device="cuda")
    call("p2pb_voxel_prep", p(coords), B, N, r, 1, ctypes.c_float(0.0), p(nc), p(ind), p(order), p(start), p(cnt), s)
    ref = dense.alloc_padded(B, Cp, r, "cuda")
    call("p2pb_voxelize_padded", p(feat), 64, Cf, p(temb), E, p(order), p(start), p(cnt), p(ref), Cp, B, N, r, s)
    out = dense.alloc_padded(B, Cp, r, "cuda")
    args = (p(feat), 64, Cf, p(temb), E, p(order), p(ind), p(start), p(cnt), p(out), Cp, B, N, r)
    call("p2pb_voxelize_padded_sparse", *args, 0, s)
    assert torch.equal(out, ref) and int((cnt > 0).sum()) > 0
    call("p2pb_voxelize_padded_sparse", *args, 1, s)
    assert int(out.count_nonzero()) == 0
    outh = dense.alloc_padded(B, Cp, r, "cuda", torch.float16)       # half grid: same sums, rounded once
    argsh = args[:9] + (p(outh),) + args[10:]
    call("p2pb_voxelize_padded_sparse_f16", *argsh, 0, s)
    assert torch.equal(outh, ref.half())
    call("p2pb_voxelize_padded_sparse_f16", *argsh, 1, s)
    assert int(outh.count_nonzero()) == 0


def test_dual_chain_engine_equals_single_chain(monkeypatch):
    """With engine.OPTIONS.chains = 2 a batch runs as two half-batch chains on two streams in one CUDA graph (DualEngine).
    Patches never interact, every kernel is deterministic and per-sample, so the result must equal the single-chain
    engine bit for bit, and graph replays must be reproducible."""
    from p2pb_b200.engine import DualEngine, Engine

    cfg = load_cfg("PVDS_PUNet")
    x = patch_input(16, 1024, seed=21).cuda()
    model, _ = build(cfg, backend="engine", head_scale=0.02)
    from p2pb_b200 import engine as E

    monkeypatch.setattr(E.OPTIONS, "chains", 2)
    out_dual = model.sample(x_start=x, steps=3, log_count=3, verbose=False)["x_chain"].clone()
    assert isinstance(model.last_engine, DualEngine)
    again = model.sample(x_start=x, steps=3, log_count=3, verbose=False)["x_chain"]
    assert torch.equal(out_dual, again)
    monkeypatch.setattr(E.OPTIONS, "chains", 1)
    out_single = model.sample(x_start=x, steps=3, log_count=3, verbose=False)["x_chain"]
    assert isinstance(model.last_engine, Engine)
    assert torch.equal(out_dual, out_single), (out_dual - out_single).abs().max()


def test_engine_tf32_operand_path_still_matches_golden(golden_dir, monkeypatch):
    """engine.OPTIONS.halo_f16 = gemm_f16 = False: every contraction with fp32-stored (tf32) operands -- the path the IEEE-half
    operand storage replaced -- stays available and inside the same tolerance.  Measured on B200 vs the fp32 golden:
    half operands mean|err| 4.2e-4 / max 5.3e-3, tf32 operands 3.9e-4 / 5.6e-3; between the two 3.4e-4 / 5.5e-3 (both round
    the operands to 10 mantissa bits, half to nearest, tf32 by truncation)."""
    from p2pb_b200 import engine as E
    from p2pb_b200.engine import get_engine

    z, cfg = _golden(golden_dir, "pvds_b2")
    x = torch.from_numpy(z["x_start"]).cuda()
    B, _, N = x.shape
    nl = float(z["noise_level"][0])
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setattr(E.OPTIONS, "halo_f16", flag == "1")
        monkeypatch.setattr(E.OPTIONS, "gemm_f16", flag == "1")
        model, _ = build(cfg, backend="engine")
        eng = get_engine(model, model.model, x.shape, None)
        assert eng.halo_f16 == (flag == "1")
        with torch.no_grad():
            sin = eng.time_embedding(nl, None)[None].expand(B, -1).contiguous()
            eps = eng.evaluate(x.contiguous(), sin)[:, :3].reshape(B, N, 3).permute(0, 2, 1).cpu().numpy()
        err = np.abs(eps - z["eps"])
        print(f"operands {'half' if flag == '1' else 'tf32'}: mean|err|={err.mean():.3e} max|err|={err.max():.3e}")
        assert err.mean() <= 2e-3 and err.max() <= 3e-2
        outs[flag] = eps
    d = np.abs(outs["1"] - outs["0"])
    print(f"half vs tf32 operands: mean|diff|={d.mean():.3e} max|diff|={d.max():.3e}")
    assert d.mean() <= 1e-3


def test_sample_with_two_step_counts_and_reloaded_weights():
    """One model, sample(steps=3) then sample(steps=5) (the reference takes `steps` per call, p2pb.py:337-363), then a
    load_state_dict: the engine must be rebuilt from the new weights (packed weights are a snapshot), and reloading the old
    weights must reproduce the first result bit for bit."""
    from p2pb_b200.model_loader import seeded_state_dict

    cfg = load_cfg("PVDS_PUNet")
    model, _ = build(cfg, backend="engine", head_scale=0.02)
    x = patch_input(2, 1024, seed=3).cuda()
    a3 = model.sample(x_start=x, steps=3, log_count=3, verbose=False)["x_pred"].clone()
    a5 = model.sample(x_start=x, steps=5, log_count=5, verbose=False)["x_pred"].clone()
    assert not torch.equal(a3, a5)
    assert torch.equal(a3, model.sample(x_start=x, steps=3, log_count=3, verbose=False)["x_pred"])
    sd0 = {k: v.clone() for k, v in model.model.state_dict().items()}
    model.model.load_state_dict(seeded_state_dict(model.model, seed=1, head_scale=0.02))
    b3 = model.sample(x_start=x, steps=3, log_count=3, verbose=False)["x_pred"].clone()
    assert not torch.equal(a3, b3), "engine kept the packed weights of the previous state dict"
    model.model.load_state_dict(sd0)
    assert torch.equal(a3, model.sample(x_start=x, steps=3, log_count=3, verbose=False)["x_pred"])


def test_engine_refuses_configs_it_does_not_implement():
    """No silent dispatch to the library-layer path: a stochastic (ot_ode=false) sampler raises with a clear message."""
    cfg = load_cfg("PVDS_PUNet", **{"diffusion.ot_ode": False})
    model, _ = build(cfg, backend="engine", head_scale=0.02)
    x = patch_input(1, 1024, seed=3).cuda()
    with pytest.raises(NotImplementedError, match="ot_ode"):
        model.sample(x_start=x, steps=2, log_count=1, verbose=False)


def _scaled_case(scale_keys, factor):
    """PVDS, 1 x 1024 points, one seeded checkpoint with some tensors scaled by `factor`; engine vs the CPU oracle on the SAME
    weights, T = 2 free-running (damped head)."""
    from oracle import model as OM
    from oracle import ops as OO
    from p2pb_b200.config import Config
    from p2pb_b200.p2pb import P2PB
    from p2pb_b200.unet_pvc import PVCNN2Unet

    cfg_dict = load_cfg("PVDS_PUNet")
    sd = OM.make_state_dict(cfg_dict, seed=0, head_scale=0.02)
    for k in scale_keys:
        sd[k] = sd[k] * factor
    cfg = Config.wrap(cfg_dict)
    cfg.gpu = "cuda:0"
    cfg.model.ema = False
    net = PVCNN2Unet(cfg)
    net.load_state_dict(sd, strict=True)
    model = P2PB(cfg, net.cuda()).eval()
    x = patch_input(1, 1024, seed=2)
    ref = OM.sample(sd, cfg_dict, x, None, steps=2, log_count=2)["x_pred"]
    return model, x.cuda(), ref, OO


def test_half_operand_guard_weights_outside_half_range_fall_back_per_layer():
    """A checkpoint whose first voxel convolution has weights of magnitude 3e5 (> 65504, the largest finite IEEE half): stored as
    half they would be +-inf.  The engine must keep THAT layer in fp32 / tf32 storage (the reference's arithmetic), say so, and
    still match the oracle; every other layer keeps half operands."""
    from p2pb_b200 import ops

    keys = ["sa_layers.0.0.voxel_layers.0.weight", "sa_layers.0.0.voxel_layers.0.bias"]
    model, x, ref, _ = _scaled_case(keys, 1e7)
    out = model.sample(x_start=x, steps=2, log_count=2, verbose=False)["x_pred"]
    eng = model.last_engine
    assert any("sa_layers.0.0.voxel_layers" in t for t in eng.tf32_layers), eng.tf32_layers
    assert len(eng.tf32_layers) == 1 and eng.halo_f16 and eng.gemm_f16
    assert torch.isfinite(out).all()
    cd = ops.calculate_cd(out, ref.cuda())
    d = (out.cpu() - ref).abs()
    print(f"weights x1e7 in one conv: tf32 layers = {eng.tf32_layers}; chamfer vs oracle {max(cd):.2e}, mean|diff| {d.mean():.2e}")
    assert max(cd) < 1e-5 and d.mean().item() < 1e-3
    # tiny weights (sub-normal in half) fall back too
    model2, x2, ref2, _ = _scaled_case(["fp_layers.3.1.voxel_layers.4.weight"], 1e-5)
    out2 = model2.sample(x_start=x2, steps=2, log_count=2, verbose=False)["x_pred"]
    assert any("fp_layers.3.1.voxel_layers" in t for t in model2.last_engine.tf32_layers)
    assert max(ops.calculate_cd(out2, ref2.cuda())) < 1e-5


def test_half_operand_guard_activation_overflow_switches_to_tf32_storage():
    """GroupNorm gains of 1e6 make the activations between the two voxel convolutions ~1e6: they do not fit half.  The
    half-producing kernels count the overflow, P2PB.sample warns, rebuilds the engine with fp32 / tf32 storage and repeats the
    call; the result matches the oracle and later calls stay on the safe path."""
    from p2pb_b200 import ops

    model, x, ref, _ = _scaled_case(["sa_layers.0.0.voxel_layers.1.norm.weight"], 1e6)
    with pytest.warns(UserWarning, match="IEEE-half range"):
        out = model.sample(x_start=x, steps=2, log_count=2, verbose=False)["x_pred"]
    eng = model.last_engine
    assert not eng.halo_f16 and not eng.gemm_f16
    assert torch.isfinite(out).all()
    cd = ops.calculate_cd(out, ref.cuda())
    d = (out.cpu() - ref).abs()
    print(f"activations ~1e6: chamfer vs oracle {max(cd):.2e}, mean|diff| {d.mean():.2e}")
    assert max(cd) < 1e-5 and d.mean().item() < 1e-3
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("error")
        out2 = model.sample(x_start=x, steps=2, log_count=2, verbose=False)["x_pred"]
    assert torch.equal(out, out2)


@pytest.mark.parametrize("C", [32, 64, 96, 128])
def test_group_project_matches_torch(C):
    """First set-abstraction layer in gather-after-GEMM form: v = Pf[idx] + Wx.(xyz[idx] - centre); mode 0 = per-centre GroupNorm
    partials, mode 1 = swish(v*A + B) as IEEE-half rows.  Against plain torch fp32 (C = 32 / 96: the half-used last 64-channel pass)."""
    import ctypes

    from p2pb_b200._lib import call

    vp = ctypes.c_void_p
    p = lambda t: vp(t.data_ptr()) if t is not None else vp(0)
    s = vp(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(C)
    B, N, M, K = 3, 300, 37, 32
    Pf = torch.randn(B * N, C, device="cuda", generator=g)
    Wx = torch.randn(C, 3, device="cuda", generator=g)
    coords = torch.randn(B, 3, N, device="cuda", generator=g)
    centers = torch.randn(B, 3, M, device="cuda", generator=g)
    idx = torch.randint(0, N, (B, M, K), device="cuda", generator=g, dtype=torch.int32)
    A = torch.randn(B, C, device="cuda", generator=g)
    Bc = torch.randn(B, C, device="cuda", generator=g)
    bi = torch.arange(B, device="cuda")[:, None, None]
    pf = Pf.view(B, N, C)[bi, idx.long()]                                             # [B,M,K,C]
    rel = coords.permute(0, 2, 1)[bi, idx.long()] - centers.permute(0, 2, 1)[:, :, None, :]   # [B,M,K,3]
    v = pf + rel @ Wx.t()
    stats = torch.zeros(B * M, C, 2, device="cuda")
    call("p2pb_group_project", p(Pf), C, p(Wx), p(coords), p(centers), p(idx), vp(0), vp(0), p(stats), vp(0), 0, B, C, N, M, K, 0, s)
    assert torch.allclose(stats[..., 0].view(B, M, C), v.sum(2), atol=1e-3, rtol=1e-4)
    assert torch.allclose(stats[..., 1].view(B, M, C), (v * v).sum(2), atol=1e-3, rtol=1e-4)
    ldo = (C + 63) // 64 * 64
    out = torch.zeros(B * M * K, ldo, device="cuda", dtype=torch.float16)
    call("p2pb_group_project", p(Pf), C, p(Wx), p(coords), p(centers), p(idx), p(A), p(Bc), vp(0), p(out), ldo, B, C, N, M, K, 1, s)
    ref = torch.nn.functional.silu(v * A[:, None, None, :] + Bc[:, None, None, :]).reshape(B * M * K, C)
    assert torch.allclose(out[:, :C].float(), ref, atol=2e-3, rtol=2e-3)
    assert int(out[:, C:].count_nonzero()) == 0
