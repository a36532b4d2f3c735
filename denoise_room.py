#!/usr/bin/env python
"""Room denoising entry point: same CLI / ``opt.yaml`` discovery / output naming / ``.ply`` output as the reference's
``denoise_room.py`` (flags :292-312, flow :424-570), running the B200 hot path.  Patch creation, normalisation and reassembly
run on the device (``p2pb_b200/room.py``, ``csrc/room.cu``); launched with ``torchrun`` (one process per GPU) every rank
creates and denoises its own shard of the patch jobs and the per-point running mean of the reference (:262-289) becomes
per-rank fixed-point sums + ONE ``all_reduce`` (NCCL over NVLink).

All reference flags are honoured: ``--average_predictions False`` (FPS of all denoised patches, :523-531, 552-556),
``--intermediate`` (one ``*_step_i.ply`` per sampling step, :475-478, 512-521, 566-570), the fill of points no patch touched
(:540-550).  ``--strict_ref`` additionally reproduces the reference's dropped last patch per chunk (:492-505) and draws the
padding randoms from np.random in its order (the default is a counter-based device RNG, rank-count independent).
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch
import torch.distributed as dist

from p2pb_b200 import room as R
from p2pb_b200.config import load_yaml
from p2pb_b200.io_ply import read_ply, write_ply
from p2pb_b200.model_loader import load_diffusion, logger


def _bool(v):
    """argparse ``type=bool`` of the reference (any non-empty string is True, denoise_room.py:298,303) -- plus the literal
    spellings of False, so that ``--average_predictions False`` does what it says."""
    return bool(v) and str(v).lower() not in ("false", "0", "no")


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--room_path", type=str, required=True, help="Path to the room point cloud.")
    p.add_argument("--model_path", type=str, required=True, help="Path to the model.")
    p.add_argument("--seed", type=int, default=42, help="Random seed.")
    p.add_argument("--use_ema", type=_bool, default=True, help="Use EMA model for prediction.")
    p.add_argument("--feature_name", type=str, default="dino_iphone")
    p.add_argument("--out_path", type=str, default=None, help="Path to save the denoised room.")
    p.add_argument("--overwrite", action="store_true", help="Overwrite existing predictions.")
    p.add_argument("--average_predictions", type=_bool, default=True, help="Average out predictions.")
    p.add_argument("--steps", type=int, default=5, help="Number of steps for the diffusion.")
    p.add_argument("--k", type=int, default=4, help="Number of patches to sample.")
    p.add_argument("--intermediate", action="store_true", help="Save intermediate steps.")
    p.add_argument("--batch_size", type=int, default=32, help="Batch size for denoising.")
    p.add_argument("--local_rank", type=int, default=int(os.environ.get("LOCAL_RANK", 0)), help="Local rank.")
    p.add_argument("--gpu", type=str, default=None, help="GPU to use.")
    p.add_argument("--distribution_type", default="none")
    p.add_argument("--strict_ref", action="store_true", help="reproduce the reference's dropped last patch per chunk and np.random padding")
    args = p.parse_args(argv)
    cfg = load_yaml(os.path.join(os.path.dirname(args.model_path), "opt.yaml"))
    cfg.merge(vars(args))
    if cfg.gpu is None:
        cfg.gpu = f"cuda:{cfg.local_rank}"
    cfg.restart = False
    return cfg


def default_out_path(cfg) -> str:
    """denoise_room.py:430-445."""
    model_training_steps = cfg.model_path.split("_")[-1].split(".")[0]
    model_config = cfg.model_path.split("/")[-2]
    ema = "_ema" if cfg.use_ema else ""
    room_source = cfg.room_path.split("/")[-1].split(".")[0]
    return os.path.join(os.path.dirname(cfg.room_path), "..", "predictions", "P2SB",
                        f"{model_config.replace('_', '-')}_{room_source.replace('_', '-')}_{model_training_steps}_{cfg.steps}{ema}.ply")


def load_room_files(cfg):
    """denoise_room.py:324-349."""
    room_points, room_colors = read_ply(cfg.room_path)
    if room_colors is not None and len(room_colors) != len(room_points):
        logger.warning("Color array has different length than point array. Setting colors to None.")
        room_colors = None
    room_dino = None
    if cfg.data.get("point_features") == "dino":
        fp = os.path.join(os.path.dirname(cfg.room_path), "..", "features", f"{cfg.feature_name}.npy")
        try:
            room_dino = np.load(fp)
            if "arkit" not in str(cfg.data.dataset).lower():
                room_dino = room_dino.T
        except Exception:
            logger.warning("No dino features found")
    return room_points, room_colors, room_dino


def main(argv=None):
    cfg = parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(cfg.gpu)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    torch.manual_seed(cfg.seed)
    np.random.seed(cfg.seed)
    out_path = os.path.abspath(cfg.out_path) if cfg.out_path else default_out_path(cfg)
    if os.path.exists(out_path) and not cfg.overwrite:
        logger.info(f"Prediction already exists at {out_path}")
        return
    model, _ = load_diffusion(cfg)
    room_points, room_colors, room_dino = load_room_files(cfg)
    dev = torch.device(cfg.gpu)
    room = torch.from_numpy(np.ascontiguousarray(room_points)).float().to(dev).contiguous()
    feats = None                                   # x_cond channels in the reference's order: rgb, then dino (:149-153)
    if cfg.data.get("use_rgb_features") and room_colors is not None:
        rgb = room_colors.astype(np.float32) / (255.0 if room_colors.dtype == np.uint8 else 1.0)
        feats = torch.from_numpy(rgb).float().to(dev)
    if cfg.data.get("point_features") == "dino" and room_dino is not None:
        d = torch.from_numpy(np.ascontiguousarray(room_dino)).float().to(dev)
        feats = d if feats is None else torch.cat([feats, d], dim=1)
    radius = 0.3 if "scannet" in str(cfg.data.dataset).lower() else 0.5
    logger.info(f"Detected dataset: {cfg.data.dataset}, denoising in radius {radius}")
    res = R.sweep(model, room, int(cfg.data.npoints), int(cfg.k), radius, int(cfg.steps), int(cfg.batch_size), int(cfg.seed),
                  feats=None if feats is None else feats.contiguous(), use_ema=bool(cfg.use_ema),
                  average_predictions=bool(cfg.average_predictions), intermediate=bool(cfg.intermediate),
                  strict_ref=bool(cfg.strict_ref), rank=rank, world=world)
    if rank == 0:
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        if cfg.average_predictions:
            out = res.denoised.cpu().numpy()
            cnt = res.count.cpu().numpy()
            missing = R.fill_not_updated(out, cnt)
            if missing:
                logger.warning(f"There are {missing} points that did not get updated.")
            write_ply(out_path, out, room_colors)
            logger.info(f"wrote {out_path} ({int((cnt > 0).sum())} of {len(cnt)} points updated, {res.n_jobs} patch jobs)")
            if res.steps is not None:
                for i, st in enumerate(res.steps):
                    s = st.cpu().numpy()
                    R.fill_not_updated(s, cnt)
                    # the reference names these f"{out_path.split('.')[0]}_step_{i}.ply" (:569), which cuts at the FIRST dot of the path
                    # -- inside the "/../" of its own default out_path; the intended stem is used here
                    write_ply(f"{os.path.splitext(out_path)[0]}_step_{i}.ply", s, room_colors)
        else:
            write_ply(out_path, res.denoised.cpu().numpy().astype(np.float64), room_colors)
            logger.info(f"wrote {out_path} (FPS of {res.n_jobs} denoised patches, no averaging)")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
