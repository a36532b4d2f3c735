#!/usr/bin/env python
"""BASELINE config 5 (and 3/4 with --extra 0/3): full-room patch sweep on N GPUs -- patch planning, device patch creation, T=30
PVDL sampling of every patch, fixed-point reassembly, ONE all_reduce -- on the synthetic room of BASELINE.md §2: the 6 walls of a
6 x 4 x 3 m box + 3 interior boxes, 2 M points, N(0, 0.01^2 m) noise (seed 0); conditioning = uniform RGB (seed 1) + 384
N(0,1) "DINOv2" channels (seed 2).  Seeded random-init PVDL weights (data.npoints = 8192).

    python tools/room_sweep.py [--points 2000000] [--extra 387] [--steps 30] [--k 4] [--batch 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/room_sweep.py ...

Every rank synthesises the same room (seeded), plans all jobs, then creates / denoises only its shard.  Timed with CUDA events
between barriers, max over ranks; rank 0 prints ONE JSON line.  The timed region starts with the room resident in HBM and ends
with the reassembled room on rank 0's device (file I/O excluded, as for the reference's metric)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_room(n_points: int, seed: int = 0):
    """Points uniform (by area) on the 6 inner faces of a 6x4x3 box and the faces of 3 interior boxes, + N(0, 0.01^2) noise."""
    import numpy as np

    rng = np.random.default_rng(seed)
    boxes = [((0, 0, 0), (6, 4, 3)), ((1.0, 1.0, 0), (2.0, 2.2, 0.9)), ((3.5, 0.5, 0), (5.0, 1.3, 1.6)), ((2.5, 2.8, 0), (3.3, 3.6, 2.1))]
    faces = []
    for lo, hi in boxes:
        lo, hi = np.array(lo, float), np.array(hi, float)
        for ax in range(3):
            for v in (lo[ax], hi[ax]):
                o = [a for a in range(3) if a != ax]
                faces.append((ax, v, lo, hi, (hi[o[0]] - lo[o[0]]) * (hi[o[1]] - lo[o[1]])))
    area = np.array([f[4] for f in faces])
    which = rng.choice(len(faces), n_points, p=area / area.sum())
    pts = np.empty((n_points, 3))
    for i, (ax, v, lo, hi, _) in enumerate(faces):
        m = which == i
        p = rng.uniform(lo, hi, size=(int(m.sum()), 3))
        p[:, ax] = v
        pts[m] = p
    return (pts + rng.normal(0, 0.01, pts.shape)).astype(np.float32), float(area.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=2_000_000)
    ap.add_argument("--extra", type=int, default=387, help="x_cond channels: 0 (config 3), 3 (config 4), 387 (config 5)")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--npoints", type=int, default=8192)
    ap.add_argument("--radius", type=float, default=0.5)
    ap.add_argument("--repeat", type=int, default=1, help="timed sweeps after one warm-up sweep of 2 batches per rank")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist
    import yaml

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from p2pb_b200 import _lib
    from p2pb_b200 import room as R
    from p2pb_b200.config import Config
    from p2pb_b200.model_loader import seeded_state_dict
    from p2pb_b200.p2pb import P2PB
    from p2pb_b200.unet_pvc import PVCNN2Unet

    cfg_dict = yaml.safe_load(open(os.path.join(ROOT, "p2pb_b200", "configs", "PVDL_SNPP.yaml")))
    cfg_dict["data"]["npoints"] = args.npoints
    cfg_dict["model"]["extra_feature_channels"] = args.extra
    cfg = Config.wrap(cfg_dict)
    cfg.gpu = str(dev)
    cfg.model.ema = False
    net = PVCNN2Unet(cfg)
    net.load_state_dict(seeded_state_dict(net, seed=0, head_scale=0.02), strict=True)
    model = P2PB(cfg, net.to(dev)).eval()

    t0 = time.time()
    pts, area = synth_room(args.points, seed=0)
    room = torch.from_numpy(pts).to(dev)
    feats = None
    if args.extra:
        g1 = torch.Generator(device=dev).manual_seed(1)
        feats = torch.rand(args.points, 3, device=dev, generator=g1)
        if args.extra > 3:
            g2 = torch.Generator(device=dev).manual_seed(2)
            feats = torch.cat([feats, torch.randn(args.points, args.extra - 3, device=dev, generator=g2)], 1).contiguous()
    t_synth = time.time() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: engine build + graph capture for this rank's batch shape (planning is cheap and deterministic; the sweep re-plans)
    plan0 = R.plan_jobs(room, args.npoints, args.k, args.radius, args.batch)
    lo0, hi0 = R.shard_range(len(plan0.job_patch), rank, world)
    bs = R.balanced_batch(hi0 - lo0, args.batch)
    gw = torch.Generator(device=dev).manual_seed(3)
    xw = torch.randn(bs, 3, args.npoints, device=dev, generator=gw)
    xw = xw / xw.norm(dim=1).amax(dim=1)[:, None, None]
    cw = None if feats is None else torch.randn(bs, args.extra, args.npoints, device=dev, generator=gw)
    for _ in range(2):
        model.sample(x_start=xw, x_cond=cw, verbose=False, steps=args.steps, use_ema=False, log_count=1)
    del xw, cw, plan0
    barrier()
    best = None
    for _ in range(args.repeat):
        l0 = _lib.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        barrier()
        ev[0].record()
        plan = R.plan_jobs(room, args.npoints, args.k, args.radius, args.batch)
        J = len(plan.job_patch)
        lo, hi = R.shard_range(J, rank, world)
        ev[1].record()
        # the sweep proper (re-plans inside: planning is counted twice in `total`, once in `plan`)
        res = R.sweep(model, room, args.npoints, args.k, args.radius, args.steps, args.batch, 42, feats=feats, rank=rank, world=world)
        ev[2].record()
        barrier()
        ms_plan, ms_sweep = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        t = torch.tensor([ms_plan, ms_sweep], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_plan, ms_sweep = float(t[0]), float(t[1])
        if best is None or ms_sweep < best[1]:
            best = (ms_plan, ms_sweep, J, hi - lo, _lib.launch_count() - l0, res)
    ms_plan, ms_sweep, J, nj, launches, res = best
    if rank == 0:
        cnt = res.count
        den = res.denoised
        moved = (den - room.double()).norm(dim=1)
        line = {
            "metric": "full-room sweep: denoised patches/sec incl. patch creation + reassembly (PVDL, N=8192, T=%d)" % args.steps,
            "value": J / (ms_sweep / 1e3), "unit": "patches/s", "n_gpus": world, "ms_sweep": ms_sweep, "ms_plan_only": ms_plan,
            "room_points": args.points, "room_points_per_s": args.points / (ms_sweep / 1e3), "patch_jobs": J, "jobs_rank0": nj,
            "x_cond_channels": args.extra, "k": args.k, "radius": args.radius, "batch_per_gpu": args.batch, "balanced_batch_rank0": bs,
            "points_updated": int((cnt > 0).sum()), "mean_updates_per_point": float(cnt.float().mean()),
            "mean_displacement_m": float(moved[cnt > 0].mean()), "surface_m2": area, "gpu_launches_rank0": int(launches),
            "synth_upload_s": t_synth, "scaling": "strong (one room, jobs sharded over ranks, one all_reduce)",
            "config": {"workload": "BASELINE configs[%d]: PVDL_SNPP npoints=8192 x_cond=%d ch, synthetic 6x4x3 m room" %
                                   (4 if args.extra > 3 else (3 if args.extra else 2), args.extra)},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
