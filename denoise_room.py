#!/usr/bin/env python
"""Room denoising entry point: same CLI / ``opt.yaml`` discovery / ``.ply`` output as the reference's ``denoise_room.py``
(flags :292-312, flow :424-570), running the B200 hot path; patches shard data-parallel across GPUs when launched
with ``torchrun`` (one process per GPU), the per-point running mean of the reference (:262-289) becomes per-rank
sums + ONE ``all_reduce`` (NCCL over NVLink) -- see ``p2pb_b200/parallel.py``.

radius patches (device radius query, ``ops.radius_query``; the reference builds a CPU KD-tree, :454-465) -> pad with
jittered duplicates / FPS down to ``npoints`` -> batched
``P2PB.sample`` -> reassembly -> ``.ply``.  Deviations from the reference, switchable with ``--strict_ref``: the
reference drops the last patch of every chunk (:498-505) -- here every patch is denoised; ``fpsample`` (absent) is
replaced by this repo's FPS kernel (start index 0).
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch
import torch.distributed as dist

from p2pb_b200 import ops
from p2pb_b200.config import load_yaml
from p2pb_b200.io_ply import read_ply, write_ply
from p2pb_b200.model_loader import load_diffusion, logger
from p2pb_b200.parallel import RoomAccumulator, shard_range


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--room_path", type=str, required=True, help="Path to the room point cloud.")
    p.add_argument("--model_path", type=str, required=True, help="Path to the model.")
    p.add_argument("--seed", type=int, default=42, help="Random seed.")
    p.add_argument("--use_ema", type=bool, default=True, help="Use EMA model for prediction.")
    p.add_argument("--feature_name", type=str, default="dino_iphone")
    p.add_argument("--out_path", type=str, default=None, help="Path to save the denoised room.")
    p.add_argument("--overwrite", action="store_true", help="Overwrite existing predictions.")
    p.add_argument("--average_predictions", type=bool, default=True, help="Average out predictions.")
    p.add_argument("--steps", type=int, default=5, help="Number of steps for the diffusion.")
    p.add_argument("--k", type=int, default=4, help="Number of patches to sample.")
    p.add_argument("--intermediate", action="store_true", help="Save intermediate steps.")
    p.add_argument("--batch_size", type=int, default=32, help="Batch size for denoising.")
    p.add_argument("--local_rank", type=int, default=int(os.environ.get("LOCAL_RANK", 0)), help="Local rank.")
    p.add_argument("--gpu", type=str, default=None, help="GPU to use.")
    p.add_argument("--distribution_type", default="none")
    p.add_argument("--strict_ref", action="store_true", help="reproduce the reference's dropped last patch per chunk")
    args = p.parse_args(argv)
    cfg = load_yaml(os.path.join(os.path.dirname(args.model_path), "opt.yaml"))
    cfg.merge(vars(args))
    if cfg.gpu is None:
        cfg.gpu = f"cuda:{cfg.local_rank}"
    cfg.restart = False
    return cfg


def create_patches(room_points, patch_size, idx_lists, room_colors=None, room_dino=None, device="cuda"):
    """denoise_room.py:352-421 -> (xyz [P,n,3], rgb, dino, idx [P,n], cut [P])."""
    xyz, rgb, dino, idxs, cuts = [], [], [], [], []
    for mapping in idx_lists:
        pts = room_points[mapping]
        n = len(pts)
        if n == 0:
            continue
        if n < patch_size:                                   # pad with jittered random duplicates (:369-395)
            extra = np.random.randint(0, n, patch_size - n)
            noise = np.linalg.norm(pts.max(0) - pts.min(0)) * 1e-2
            add = pts[extra] + np.random.normal(0, noise, (patch_size - n, 3))
            sel = np.concatenate([np.arange(n), extra])
            xyz.append(np.concatenate([pts, add], 0))
            idxs.append(mapping[sel])
            cuts.append(n)
            if room_colors is not None:
                rgb.append(room_colors[mapping][sel])
            if room_dino is not None:
                dino.append(room_dino[mapping][sel])
        else:                                                # FPS down to patch_size (:400-419), n//patch_size+1 times
            c = torch.from_numpy(pts.T.copy()).float().unsqueeze(0).to(device)
            sel = ops.furthest_point_sampling(c, patch_size)[0].long().cpu().numpy()
            for _ in range(n // patch_size + 1):
                xyz.append(pts[sel])
                idxs.append(mapping[sel])
                cuts.append(patch_size)
                if room_colors is not None:
                    rgb.append(room_colors[mapping][sel])
                if room_dino is not None:
                    dino.append(room_dino[mapping][sel])
    st = lambda l: np.stack(l) if l else None
    return st(xyz), st(rgb), st(dino), st(idxs), np.array(cuts)


@torch.no_grad()
def denoise_patch_batch(xyz, model, cfg, rgb=None, dino=None):
    """denoise_room.py:115-178: per-patch centre / max-norm scale, sample, de-normalise."""
    x = torch.from_numpy(xyz).float().to(cfg.gpu)
    center = x.mean(dim=1, keepdim=True)
    x = x - center
    scale = x.norm(dim=2).amax(dim=1)[:, None, None]
    x = (x / scale).transpose(1, 2).contiguous()
    cond = None
    if cfg.data.get("use_rgb_features") and rgb is not None:
        cond = torch.from_numpy(rgb).float().to(cfg.gpu).transpose(1, 2)
    if cfg.data.get("point_features") == "dino" and dino is not None:
        d = torch.from_numpy(dino).float().to(cfg.gpu).transpose(1, 2)
        cond = d if cond is None else torch.cat([cond, d], dim=1)
    out = model.sample(x_start=x, x_cond=None if cond is None else cond.contiguous(), verbose=False, steps=cfg.steps,
                       use_ema=cfg.use_ema, log_count=1)["x_pred"]
    return out.transpose(1, 2) * scale + center


def main(argv=None):
    cfg = parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(cfg.local_rank)
        dist.init_process_group("nccl")
    torch.manual_seed(cfg.seed)
    np.random.seed(cfg.seed)          # same seed on every rank: identical patch lists, then sharded
    out_path = os.path.abspath(cfg.out_path) if cfg.out_path else os.path.join(
        os.path.dirname(cfg.room_path), "..", "predictions", "P2SB", os.path.basename(cfg.room_path))
    if os.path.exists(out_path) and not cfg.overwrite:
        logger.info(f"Prediction already exists at {out_path}")
        return
    model, _ = load_diffusion(cfg)
    room_points, room_colors = read_ply(cfg.room_path)
    room_dino = None
    if cfg.data.get("point_features") == "dino":
        fp = os.path.join(os.path.dirname(cfg.room_path), "..", "features", f"{cfg.feature_name}.npy")
        if os.path.exists(fp):
            room_dino = np.load(fp)
            if "arkit" not in str(cfg.data.dataset).lower():
                room_dino = room_dino.T
    npts = cfg.data.npoints
    n_centers = int(np.ceil(room_points.shape[0] / npts) * cfg.k)
    pts_dev = torch.from_numpy(room_points).float().to(cfg.gpu).contiguous()
    c = pts_dev.t().contiguous().unsqueeze(0)
    center_idx = ops.furthest_point_sampling(c, n_centers)[0].long()
    radius = 0.3 if "scannet" in str(cfg.data.dataset).lower() else 0.5
    # KDTree.query_radius of the reference (denoise_room.py:454-465) on the device: CSR with ascending indices per centre
    off, idx = ops.radius_query(pts_dev[center_idx].contiguous(), pts_dev, radius)
    off, idx = off.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
    idx_lists = [idx[off[i]:off[i + 1]] for i in range(n_centers)]
    xyz, rgb, dino, idxs, cuts = create_patches(room_points, npts, idx_lists, room_colors, room_dino, cfg.gpu)
    P = xyz.shape[0]
    lo, hi = shard_range(P, rank, world)
    acc = RoomAccumulator(room_points.shape[0], device=cfg.gpu)
    bs = cfg.batch_size
    for s in range(lo, hi, bs):
        e = min(s + bs, hi)
        sl = np.arange(s, e)
        if cfg.strict_ref and len(sl) > 1:
            sl = sl[:-1]                                     # the reference's [start:end] with end = last index
        pad = bs - len(sl)                                   # static batch shape for the captured graph
        take = np.concatenate([sl, np.repeat(sl[-1:], pad)]) if pad else sl
        den = denoise_patch_batch(xyz[take], model, cfg, None if rgb is None else rgb[take], None if dino is None else dino[take])
        for j, p in enumerate(sl):
            acc.add(torch.from_numpy(idxs[p][: cuts[p]]), den[j, : cuts[p]])
    mean, count = acc.reduce()
    if rank == 0:
        out = room_points.copy()
        m = count.cpu().numpy() > 0
        out[m] = mean.cpu().numpy()[m]
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        write_ply(out_path, out, room_colors)
        logger.info(f"wrote {out_path} ({int(m.sum())} of {len(m)} points updated)")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
