#!/usr/bin/env python
"""Object denoising entry point: same CLI, same ``opt.yaml`` + ``.pth`` discovery and same output format as the
reference's ``denoise_object.py`` (flags :19-30, flow :125-170), running the B200 hot path.

normalise -> FPS seeds -> kNN-2048 patches -> P2PB.sample (fused engine) -> de-normalise -> FPS merge -> ``.xyz``.
The seed / merge FPS (``torch_cluster.fps`` in the reference, start index 0) and the patch kNN (``pytorch3d.knn_points``,
ascending distance) run on this repo's kernels: FPS and a radix-select + sort kNN (csrc/metrics.cu).
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from p2pb_b200 import ops
from p2pb_b200.config import Config, load_yaml
from p2pb_b200.io_ply import read_ply, write_array_to_xyz
from p2pb_b200.model_loader import load_diffusion, logger


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--data_path", type=str, required=True, help="Path to the point cloud.")
    p.add_argument("--save_path", type=str, required=True, help="Output root directory.")
    p.add_argument("--model_path", type=str, required=True, help="Path to the model.")
    p.add_argument("--seed", type=int, default=42, help="Random seed.")
    p.add_argument("--k", type=int, default=3, help="Patch oversampling factor.")
    p.add_argument("--use_ema", action="store_true", help="Use EMA model for prediction.")
    p.add_argument("--gpu", type=str, default="cuda:0", help="GPU to use.")
    p.add_argument("--steps", type=int, default=5, help="Number of steps for the diffusion.")
    p.add_argument("--distribution_type", default="none")
    args = p.parse_args(argv)
    cfg = load_yaml(os.path.join(os.path.dirname(args.model_path), "opt.yaml"))   # config lives next to the checkpoint
    cfg.merge(vars(args))
    cfg.restart = False
    cfg.local_rank = 0
    return cfg


def normalize_unit_sphere(pcl: torch.Tensor):
    """utils/utils.py:96-109 NormalizeUnitSphere.normalize: bbox centre, max-norm scale."""
    center = (pcl.max(dim=0, keepdim=True)[0] + pcl.min(dim=0, keepdim=True)[0]) / 2
    pcl = pcl - center
    scale = (pcl ** 2).sum(dim=1, keepdim=True).sqrt().max(dim=0, keepdim=True)[0]
    return pcl / scale, center, scale


def farthest_point_sampling(pcls: torch.Tensor, num_pnts: int):
    """models/evaluation.py:297-311 semantics (deterministic start at index 0) on the repo's FPS kernel. pcls [B,N,3]."""
    coords = pcls.transpose(1, 2).contiguous().float()
    idx = ops.furthest_point_sampling(coords, num_pnts).long()
    sampled = torch.gather(pcls, 1, idx.unsqueeze(-1).expand(-1, -1, 3))
    return sampled, [i for i in idx]


def knn_patches(seeds: torch.Tensor, pcl: torch.Tensor, K: int) -> torch.Tensor:
    """K nearest points of ``pcl [N,3]`` for every seed ``[P,3]``, ascending distance (pytorch3d.ops.knn_points order)."""
    idx = ops.knn_points(seeds.contiguous().float(), pcl.contiguous().float(), K).long()
    return pcl[idx]


@torch.no_grad()
def patch_based_denoise(model, pcl_noisy: torch.Tensor, patch_size: int, seed_k: int = 3, cfg=None):
    """denoise_object.py:64-122: one batch of P = int(seed_k * N / patch_size) patches, ONE global scale."""
    assert pcl_noisy.dim() == 2
    N, d = pcl_noisy.shape
    seeds, _ = farthest_point_sampling(pcl_noisy.unsqueeze(0), int(seed_k * N / patch_size))
    patches = knn_patches(seeds[0], pcl_noisy, patch_size)                       # [P, K, 3]
    centers = patches.mean(dim=1, keepdim=True)
    patches = patches - centers
    scale = torch.max(torch.norm(patches, dim=-1))
    patches = patches / scale
    out = model.sample(x_start=patches.transpose(1, 2).contiguous(), use_ema=cfg.use_ema, steps=cfg.steps,
                       log_count=cfg.steps, verbose=False)
    den = out["x_pred"].transpose(1, 2) * scale + centers
    merged, _ = farthest_point_sampling(den.reshape(1, -1, d), N)
    return merged[0]


def sample(cfg) -> None:
    torch.manual_seed(cfg.seed)
    np.random.seed(cfg.seed)
    torch.cuda.set_device(cfg.gpu)           # --gpu cuda:1: buffers, streams and graph capture all on that device
    model, _ = load_diffusion(cfg)
    model.eval()
    if cfg.data_path.endswith("ply"):
        pts, _ = read_ply(cfg.data_path)
        pcl = torch.tensor(pts, dtype=torch.float32)
    else:
        pcl = torch.tensor(np.loadtxt(cfg.data_path), dtype=torch.float32)
    pcl, center, scale = normalize_unit_sphere(pcl)
    den = patch_based_denoise(model, pcl.to(cfg.gpu), patch_size=2048, seed_k=cfg.k, cfg=cfg).cpu()
    den = den * scale + center
    if not cfg.data_path.endswith("xyz"):
        raise NotImplementedError("Only .xyz files are supported for now.")     # same restriction as the reference (:166-169)
    write_array_to_xyz(cfg.save_path, den.numpy())
    logger.info(f"wrote {cfg.save_path}")


if __name__ == "__main__":
    sample(parse_args())
