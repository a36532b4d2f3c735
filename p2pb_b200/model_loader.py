"""``load_model`` / ``load_diffusion`` / ``extract_from_state_dict`` with the reference's signatures and checkpoint
format (``models/model_loader.py:64-179``): ``{"step", "model_state", "optimizer_state"}`` where ``model_state`` is
``P2PB.state_dict()`` (keys ``model.<backbone>`` -- ``model.module.<backbone>`` when saved under DP/DDP -- and
``ema.*``), next to an ``opt.yaml``.  The optimizer/scheduler half of the reference loader is training-only and
not provided."""
from __future__ import annotations

import logging
from typing import Dict

import torch

from .p2pb import P2PB
from .unet_pvc import PVCNN2Unet

try:  # the reference logs through loguru; fall back to logging when absent
    from loguru import logger
except ImportError:  # pragma: no cover
    logger = logging.getLogger("p2pb_b200")
    logger.success = logger.info


def load_model(cfg) -> torch.nn.Module:
    model = PVCNN2Unet(cfg)
    n = sum(p.numel() for p in model.parameters() if p.requires_grad) / 1e6
    logger.info(f"Generated model with following number of params (M): {n:.2f}")
    return model


def extract_from_state_dict(state_dict: Dict, pattern: str) -> Dict:
    return {k.replace(pattern, ""): v for k, v in state_dict.items() if k.startswith(pattern)}


def load_diffusion(cfg) -> tuple:
    """Build backbone + bridge on ``cuda:<cfg.local_rank>`` and load ``cfg.model_path`` if set -> (model, ckpt)."""
    rank = cfg.get("local_rank", 0)
    device = torch.device("cuda", int(rank)) if not isinstance(rank, str) else torch.device(rank)
    if cfg.get("gpu") is None:
        cfg.gpu = str(device)
    backbone = load_model(cfg).to(cfg.gpu)
    model = P2PB(cfg=cfg, model=backbone)
    dist_type = cfg.get("distribution_type", "none")
    if dist_type in ("multi", "single"):
        # DP/DDP wrappers are a training concern; inference shards patches across processes instead.
        logger.info(f"distribution_type={dist_type}: inference runs one process per GPU, model left unwrapped")
    cfg.start_step = 0
    ckpt = None
    if cfg.get("model_path", "") not in ("", None):
        ckpt = torch.load(cfg.model_path, map_location=torch.device("cpu"), weights_only=False)
        if not cfg.get("restart", False):
            cfg.start_step = ckpt.get("step", -1) + 1
        state = ckpt["model_state"]
        # reference prefix logic (model_loader.py:125-130): "model." for multi/single else "model.module.";
        # accept either so checkpoints saved with or without a DP/DDP wrapper both load.
        model_dict = extract_from_state_dict(state, "model.module.") or extract_from_state_dict(state, "model.")
        ema_dict = extract_from_state_dict(state, "ema.")
        try:
            model.model.load_state_dict(model_dict)
            if cfg.get("use_ema", False) and ema_dict and model.ema is not None:
                ema_dict = {k.replace("ema_model.module.", "ema_model.").replace("online_model.module.", "online_model."): v
                            for k, v in ema_dict.items()}
                # ema_pytorch keeps bookkeeping buffers this minimal holder does not have (and vice versa): load non-strictly,
                # but a key mismatch INSIDE ema_model (the weights that are evaluated) is an error, not a silent skip
                res = model.ema.load_state_dict(ema_dict, strict=False)
                bad = [k for k in list(res.missing_keys) + list(res.unexpected_keys) if k.startswith("ema_model.")]
                if bad:
                    raise RuntimeError(f"EMA weights do not match the model: {bad[:8]}{' ...' if len(bad) > 8 else ''}")
                other = [k for k in list(res.missing_keys) + list(res.unexpected_keys) if not k.startswith("ema_model.")]
                if other:
                    logger.info(f"EMA bookkeeping keys not loaded: {other[:6]}{' ...' if len(other) > 6 else ''}")
                logger.success("Loaded EMA from checkpoint!")
            logger.success("Loaded Model from checkpoint!")
        except RuntimeError as e:
            logger.warning("Could not load model state dict. Trying to load without strict flag.")
            logger.warning(e)
            model.load_state_dict(state, strict=False)
        logger.info("Loaded model from %s" % cfg.model_path)
    return model, ckpt


def save_checkpoint(path: str, model: P2PB, step: int = 0) -> None:
    """Write a checkpoint in the reference's format (train.py:169-174); used to make synthetic seeded checkpoints."""
    torch.save({"step": step, "model_state": model.state_dict(), "optimizer_state": {}}, path)


def seeded_state_dict(net: torch.nn.Module, seed: int = 0, head_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights for ``net`` (no trained checkpoints exist offline): one independent RNG stream per
    state-dict key, so values do not depend on construction order.  Conv/linear weights ~ U(+-1/sqrt(fan_in))
    (AdaGN ``emd`` at half scale), norm gains 1+0.1 N(0,1), biases 0.05 N(0,1) with the AdaGN (1, 0) centre."""
    import math
    import zlib

    out = {}
    for key, ref in net.state_dict().items():
        shp = tuple(ref.shape)
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        if key.endswith(".weight") and len(shp) >= 2:
            bound = 1.0 / math.sqrt(float(torch.tensor(shp[1:]).prod()))
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
            t = t * head_scale   # < 1: damped noise head, keeps the free-running T-step map well conditioned
        out[key] = t.float()
    return out
