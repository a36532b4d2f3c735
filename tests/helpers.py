"""Shared test inputs (seeded, synthetic) and config helpers."""
import os

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG_DIR = os.path.join(ROOT, "p2pb_b200", "configs")


def cloud(B, N, seed, snap=None, scale=0.5):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, N, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * (1 + 0.05 * torch.randn(B, 1, N, generator=g))
    x = x * scale
    if snap:
        x = torch.round(x * snap) / snap
    return x.contiguous()


def patch_input(B, N, seed):
    """Same generator as oracle/gen_golden.py::make_input('synth')."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, N, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * (1 + 0.05 * torch.randn(B, 1, N, generator=g))
    x = x - x.mean(2, keepdim=True)
    return (x / x.norm(dim=1).amax(dim=1)[:, None, None]).contiguous()


def load_cfg(name, **over):
    cfg = yaml.safe_load(open(os.path.join(CFG_DIR, name + ".yaml")))
    for k, v in over.items():
        node = cfg
        ks = k.split(".")
        for kk in ks[:-1]:
            node = node[kk]
        node[ks[-1]] = v
    return cfg
