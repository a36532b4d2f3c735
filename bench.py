#!/usr/bin/env python
"""bench.py -- denoised patches/sec of the P2P-Bridge hot path (BASELINE.json metric) on N B200s of one node.

A "step" = one ``P2PB.sample`` call (T=30 bridge steps of the PVDS PVCNN U-Net) over one batch of 64 synthetic
2048-point patches per GPU (BASELINE.json configs[1]); weights are seeded random-init in the reference's checkpoint
layout (no trained weights exist offline).  Patches shard data-parallel over ranks with no data-path collective
(weak scaling: 64 patches per GPU); timing is CUDA events on the launching stream, max over ranks.

  value  patches/s with the batch already resident in HBM
  e2e    same metric through the public API with HOST buffers (pinned H2D of the batch + D2H of x_pred per step)
  roofline      dominant kernel (the implicit-GEMM voxel convolution) timed live with CUDA events after the run
  cpu_baseline  the CPU restatement (oracle/, checker only) timed on the host cores on a bounded sample (rank 0, N=1)

``--impl reference`` times the reference's CPU path for the same metric/config: the reference's ops are CUDA-only, so
its CPU implementation is the oracle port (oracle/model.py + oracle/p2pb_oracle.c, pinned against the reference's own
code and kernels), run with all host threads on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "denoised patches/sec (PVDS, N=2048, T=30)"
UNIT = "patches/s"
WORKLOAD = "PVDS_PUNet N=2048 T=30 batch=64/GPU (BASELINE configs[1])"
B_PER_GPU, NPTS, TSTEPS = 64, 2048, 30


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--backend", default="engine", choices=["engine", "eager"],
                    help="eager = the validation path (torch library layers); the bench line of record is the engine")
    ap.add_argument("--no-graph", action="store_true", help="profiling aid: enqueue kernels directly (ncu launch lists)")
    ap.add_argument("--chains", type=int, default=1, help="development aid: part-batch chains inside the graph")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=0|1",
                    help="development aid: set a boolean of p2pb_b200.engine.OPTIONS (A/B runs on one box)")
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="patches per GPU (default = the named config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` block (R-GPU arm, PVDL configs 3-5) and `parity`")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
def load_cfg_dict():
    import yaml

    return yaml.safe_load(open(os.path.join(ROOT, "p2pb_b200", "configs", "PVDS_PUNet.yaml")))


def synth_patches(B, N, seed):
    """64 kNN-like surface patches: noisy unit-sphere caps, per-patch centred, one global max-norm scale
    (denoise_object.py:97-100).  Synthetic (no PU-Net data offline)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    d = torch.randn(B, 3, 1, generator=g)
    d = d / d.norm(dim=1, keepdim=True)
    x = d + 0.35 * torch.randn(B, 3, N, generator=g)
    x = x / x.norm(dim=1, keepdim=True) + 0.02 * torch.randn(B, 3, N, generator=g)
    x = x - x.mean(dim=2, keepdim=True)
    return (x / x.norm(dim=1).max()).contiguous()


def cpu_port_rate(cfg, seconds_budget=20.0, threads=None):
    """patches/s of the CPU restatement on a bounded sample sized to ~seconds_budget of CPU work: n_patch patches x n_eval of
    the T=30 network evaluations (all 30 whenever one patch fits the budget)."""
    import torch

    from oracle import model as OM

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = OM.make_state_dict(cfg, seed=0)
    x = synth_patches(8, NPTS, seed=123)
    t0 = time.perf_counter()
    OM.sample(sd, cfg, x[:1], None, steps=1, log_count=1)      # one evaluation to size the sample (and warm caches)
    t_eval = time.perf_counter() - t0
    n_eval = int(max(1, min(TSTEPS, seconds_budget // max(t_eval, 1e-3))))
    n_patch = int(max(1, min(8, seconds_budget // max(t_eval * n_eval, 1e-3)))) if n_eval == TSTEPS else 1
    t0 = time.perf_counter()
    OM.sample(sd, cfg, x[:n_patch], None, steps=n_eval, log_count=1)
    dt = time.perf_counter() - t0
    rate = n_patch / (dt * TSTEPS / n_eval)
    return rate, threads, (f"{n_patch} patch(es) N={NPTS}, {n_eval} of T={TSTEPS} network evaluations timed ({dt:.1f} s), "
                           f"scaled to T={TSTEPS}")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def sustained_peak_tflops():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16 sustained"
    except Exception:
        return 1408.0, "fallback (BASELINE.md)"


PVDL_GFLOP_PER_EVAL = {0: 332.99, 3: 332.99, 387: 333.39}       # SURVEY.md 8(d), per patch per network evaluation at N = 8192
PVDL_B_PER_GPU, PVDL_N = 32, 8192


def pvdl_configs(dev, rank, world, timed_fn):
    """BASELINE configs 3-5 on the engine: PVDL (data.npoints = 8192), 32 patches of 8192 points per GPU, T = 30; xyz only /
    xyz+RGB (x_cond 3 ch) / xyz+RGB+DINOv2 (x_cond 387 ch).  Same timing rules as the headline (CUDA events, max over ranks,
    patches sharded over ranks, no data-path collective).  -> {name: {...}}"""
    import torch
    import yaml

    from p2pb_b200.config import Config
    from p2pb_b200.model_loader import seeded_state_dict
    from p2pb_b200.p2pb import P2PB
    from p2pb_b200.unet_pvc import PVCNN2Unet
    from tests.helpers import patch_input

    peak, peak_src = sustained_peak_tflops()
    out = {}
    for name, extra in (("pvdl_8192_xyz", 0), ("pvdl_8192_rgb", 3), ("pvdl_8192_rgb_dino", 387)):
        cfg_dict = yaml.safe_load(open(os.path.join(ROOT, "p2pb_b200", "configs", "PVDL_SNPP.yaml")))
        cfg_dict["data"]["npoints"] = PVDL_N
        cfg_dict["model"]["extra_feature_channels"] = extra
        cfg = Config.wrap(cfg_dict)
        cfg.gpu = str(dev)
        cfg.model.ema = False
        net = PVCNN2Unet(cfg)
        net.load_state_dict(seeded_state_dict(net, seed=0), strict=True)
        model = P2PB(cfg, net.to(dev)).eval()
        x = patch_input(PVDL_B_PER_GPU, PVDL_N, seed=7 + rank).to(dev)
        xc = None
        if extra:
            g = torch.Generator().manual_seed(1 if extra == 3 else 2)
            xc = torch.rand(PVDL_B_PER_GPU, 3, PVDL_N, generator=g)
            if extra > 3:
                xc = torch.cat([xc, torch.randn(PVDL_B_PER_GPU, extra - 3, PVDL_N, generator=g)], 1)
            xc = xc.to(dev)
        ms, _ = timed_fn(lambda: model.sample(x_start=x, x_cond=xc, steps=TSTEPS, log_count=1, verbose=False, use_ema=False), 2, 3)
        rate = PVDL_B_PER_GPU * world / (ms / 1e3)
        ceiling = peak * 1e12 / (PVDL_GFLOP_PER_EVAL[extra] * 1e9 * TSTEPS)
        out[name] = {"value": rate, "unit": UNIT, "ms_per_step": ms, "patches_per_gpu": PVDL_B_PER_GPU, "npoints": PVDL_N, "T": TSTEPS,
                     "x_cond_channels": extra, "algorithmic_tflops": rate * PVDL_GFLOP_PER_EVAL[extra] * TSTEPS / 1e3,
                     "ceiling_patches_s_per_gpu": ceiling, "frac_of_ceiling": rate / world / ceiling,
                     "ceiling": f"{peak_src} {peak:.1f} TFLOP/s / ({PVDL_GFLOP_PER_EVAL[extra]} GFLOP x T={TSTEPS})"}
        del model, net
        torch.cuda.empty_cache()
    return out


def rgpu_arm(local):
    """The reference's own GPU path (oracle/rgpu_bench.py) in a separate process on this rank's GPU."""
    try:
        o = subprocess.run([sys.executable, "-m", "oracle.rgpu_bench", "--device", f"cuda:{local}", "--pvdl"], cwd=ROOT,
                           capture_output=True, text=True, timeout=600)
        lines = [l for l in o.stdout.splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"available": False, "why": (o.stderr or "no output")[-300:]}
    except Exception as ex:
        return {"available": False, "why": repr(ex)[:300]}


def parity_live(dev):
    """Chamfer (calculate_cd_cuda definition) between THIS run's engine and the committed output of the unmodified reference
    on a B200 (tests/golden/rgpu_golden.npz, fp32), on the bench's own 64 patches, T = 30, damped-head seeded checkpoint --
    next to the reference's own TF32-vs-fp32 floor measured in the same golden run."""
    import numpy as np
    import torch

    from p2pb_b200 import ops
    from p2pb_b200.config import Config
    from p2pb_b200.model_loader import seeded_state_dict
    from p2pb_b200.p2pb import P2PB
    from p2pb_b200.unet_pvc import PVCNN2Unet

    z = np.load(os.path.join(ROOT, "tests", "golden", "rgpu_golden.npz"))
    cfg = Config.wrap(load_cfg_dict())
    cfg.gpu = str(dev)
    cfg.model.ema = False
    net = PVCNN2Unet(cfg)
    net.load_state_dict(seeded_state_dict(net, seed=0, head_scale=0.02), strict=True)
    model = P2PB(cfg, net.to(dev)).eval()
    x = synth_patches(B_PER_GPU, NPTS, seed=1000).to(dev)
    out = model.sample(x_start=x, steps=TSTEPS, log_count=1, verbose=False, use_ema=False)["x_pred"]
    ref = torch.from_numpy(z["cfg2_damped_x_pred_fp32"]).to(dev)
    cd = torch.tensor(ops.calculate_cd(out, ref))
    cd0 = torch.tensor(ops.calculate_cd(x, ref))
    fm, fx = float(z["floor_cd_tf32_vs_fp32_mean"]), float(z["floor_cd_tf32_vs_fp32_max"])
    return {"what": "Chamfer(engine, unmodified reference on B200 fp32), 64 bench patches, T=30, damped-head seeded checkpoint",
            "chamfer_mean": float(cd.mean()), "chamfer_max": float(cd.max()), "chamfer_do_nothing_mean": float(cd0.mean()),
            "reference_floor_tf32_vs_fp32": {"mean": fm, "max": fx},
            "reference_floor_run_to_run": {"mean": float(z["floor_cd_run_to_run_mean"]), "max": float(z["floor_cd_run_to_run_max"])},
            "bound": {"mean": max(1e-5, 2 * fm), "max": max(1e-5, 2 * fx)},
            "ok": bool(cd.mean() <= max(1e-5, 2 * fm) and cd.max() <= max(1e-5, 2 * fx))}


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: CPU implementation of the path on the host cores (oracle port; see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = load_cfg_dict()
    import torch

    from oracle import model as OM

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = OM.make_state_dict(cfg, seed=0)
    x = synth_patches(1, NPTS, seed=123)
    t0 = time.perf_counter()
    OM.sample(sd, cfg, x, None, steps=1, log_count=1)
    t_eval = time.perf_counter() - t0
    total_steps = args.steps + args.warmup
    n_eval = int(max(1, min(TSTEPS, (180.0 / max(total_steps, 1)) // max(t_eval, 1e-3))))
    for _ in range(args.warmup):
        OM.sample(sd, cfg, x, None, steps=n_eval, log_count=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        OM.sample(sd, cfg, x, None, steps=n_eval, log_count=1)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    rate = 1.0 / (dt * TSTEPS / n_eval)
    sample = f"per step: 1 patch N={NPTS}, {n_eval} of T={TSTEPS} network evaluations, scaled to T={TSTEPS}"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from p2pb_b200 import _lib
    from p2pb_b200.config import Config
    from p2pb_b200.model_loader import seeded_state_dict
    from p2pb_b200.p2pb import P2PB
    from p2pb_b200.unet_pvc import PVCNN2Unet

    from p2pb_b200 import engine as ENG

    ENG.OPTIONS.no_graph = bool(args.no_graph)
    ENG.OPTIONS.chains = int(args.chains)
    for kv in args.opt:
        k, v = kv.split("=")
        if not isinstance(getattr(ENG.OPTIONS, k, None), bool):
            raise SystemExit(f"--opt {k}: not a boolean option of engine.OPTIONS")
        setattr(ENG.OPTIONS, k, bool(int(v)))
    cfg_dict = load_cfg_dict()
    cfg = Config.wrap(cfg_dict)
    cfg.gpu = str(dev)
    cfg.model.ema = False
    cfg.backend = args.backend
    net = PVCNN2Unet(cfg)
    net.load_state_dict(seeded_state_dict(net, seed=0), strict=True)
    model = P2PB(cfg, net.to(dev)).eval()

    B = args.batch
    host_x = synth_patches(B, NPTS, seed=1000 + rank).pin_memory()          # each rank: its own shard of patches
    host_out = torch.empty((B, 3, NPTS), dtype=torch.float32).pin_memory()
    x_dev = host_x.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def step_resident():
        return model.sample(x_start=x_dev, steps=TSTEPS, log_count=1, verbose=False, use_ema=False)["x_pred"]

    def step_e2e():
        xd = host_x.to(dev, non_blocking=True)
        out = model.sample(x_start=xd, steps=TSTEPS, log_count=1, verbose=False, use_ema=False)["x_pred"]
        host_out.copy_(out, non_blocking=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            flush.zero_()            # L2 flush between timed iterations (working set is >> L2 anyway)
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, _lib.launch_count() - l0

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step, launches = timed(step_resident, args.steps, args.warmup)
    ms_e2e, _ = timed(step_e2e, max(2, args.steps // 2), 1)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    eng = getattr(model, "last_engine", None)
    if eng is not None:
        launches = eng.kernels_per_sample * args.steps
    total_patches = B * world
    value = total_patches / (ms_step / 1e3)
    e2e = total_patches / (ms_e2e / 1e3)

    roofline = None
    if rank == 0 and not args.no_roofline:
        try:
            from p2pb_b200 import roofline as RL

            roofline = RL.dominant_kernel_roofline(model, B, dev)
        except Exception as ex:  # keep the bench line even if the micro-timing fails
            roofline = {"error": repr(ex)}
    dtype_name = getattr(eng, "dtype_name", "tf32") if eng is not None else "tf32"
    extra, parity = None, None
    if not args.no_extra and args.backend == "engine":
        del model, net, eng                                   # free the headline engine's buffers before the PVDL engines
        torch.cuda.empty_cache()
        try:
            extra = pvdl_configs(dev, rank, world, timed)
        except Exception as ex:
            extra = {"error": repr(ex)[:300]}
        if rank == 0:
            try:
                parity = parity_live(dev)
            except Exception as ex:
                parity = {"error": repr(ex)[:300]}
            torch.cuda.empty_cache()
            if world == 1:
                rg = rgpu_arm(local)
                extra["rgpu"] = rg
                if rg.get("available"):
                    extra["rgpu_patches_s"] = rg["pvds_2048"]["patches_per_s"]
                    extra["speedup_vs_rgpu"] = value / rg["pvds_2048"]["patches_per_s"]
                    if "pvdl_8192_xyz" in rg and "pvdl_8192_xyz" in extra:
                        extra["pvdl_8192_xyz"]["rgpu_patches_s"] = rg["pvdl_8192_xyz"]["patches_per_s"]
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, cores, sample = cpu_port_rate(cfg_dict)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype_name, "data": "synthetic",
            "config": {"workload": WORKLOAD, "backend": args.backend, "patches_per_gpu": B, "npoints": NPTS, "T": TSTEPS,
                       "weights": "seeded random-init, reference checkpoint layout",
                       "l2": "256 MiB L2 flush between timed iterations; per-step working set >> 126 MB L2",
                       "parallelism": f"dp{world} over patches, no data-path collective"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": host_x.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4},
            "gpu_launches": int(launches),
            "clocks": sampler.summary() if sampler else None,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
