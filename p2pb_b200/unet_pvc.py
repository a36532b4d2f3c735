"""``PVCNN2Unet``: the PVCNN U-Net backbone, same constructor config, state-dict keys and
``forward(x, t, x_cond=None) -> eps`` contract as the reference (``models/unet_pvc.py:27-269``).

``forward`` is the eager channel-first path (this repo's CUDA point/voxel ops + torch dense layers);
the fused engine (``p2pb_b200.engine``) evaluates the same network from the same parameters."""
from __future__ import annotations

from functools import partial

import numpy as np
import torch
import torch.nn as nn

from .pvcnn import (Attention, LinearAttention, Pnet2Stage, PVCData, SharedMLP, Swish, create_fp_components,
                    create_mlp_components, create_pvc_layer_params, create_sa_components)


def _default(v, d):
    return v if v is not None else d


class PVCNN2Unet(nn.Module):
    def __init__(self, cfg, return_layers: bool = False):
        super().__init__()
        m, pvd = cfg.model, cfg.model.PVD
        self.return_layers = return_layers
        self.input_dim = _default(m.get("in_dim"), 3)
        self.extra_feature_channels = pvd["extra_feature_channels"] if "extra_feature_channels" in pvd else m["extra_feature_channels"]
        self.embed_dim = _default(m.get("time_embed_dim"), 64)
        out_dim = _default(m.get("out_dim"), 3)
        dropout = _default(m.get("dropout"), 0.1)
        attn_type = _default(pvd.get("attention_type"), "linear")
        self.embedf = nn.Sequential(nn.Linear(self.embed_dim, self.embed_dim), nn.LeakyReLU(0.1, inplace=True),
                                    nn.Linear(self.embed_dim, self.embed_dim))
        if pvd.get("use_global_embedding"):
            self.cond_emb_dim = c = pvd["global_embedding_dim"]
            self.global_pnet = Pnet2Stage([self.input_dim, c // 8, c // 4], [c // 2, c])
        else:
            self.global_pnet, self.cond_emb_dim = None, 0
        self.f_embed_dim = pvd.get("feat_embed_dim", self.extra_feature_channels)
        self.embed_feats = None
        if self.f_embed_dim != self.extra_feature_channels:
            cin = self.extra_feature_channels or self.input_dim
            self.embed_feats = nn.Sequential(nn.Conv1d(cin, self.f_embed_dim, 1), nn.GroupNorm(8, self.f_embed_dim), Swish(),
                                             nn.Conv1d(self.f_embed_dim, self.f_embed_dim, 1))
        sa_blocks, fp_blocks = create_pvc_layer_params(
            npoints=cfg.data.npoints, channels=pvd.channels, n_sa_blocks=pvd.n_sa_blocks, n_fp_blocks=pvd.n_fp_blocks,
            radius=pvd.radius, voxel_resolutions=pvd.voxel_resolutions, centers=pvd.get("centers"))
        self.heads = pvd.attention_heads
        if str(attn_type).lower() == "linear":
            attention_fn = partial(LinearAttention, heads=pvd.attention_heads)
        elif str(attn_type).lower() == "flash":
            attention_fn = partial(Attention, heads=pvd.attention_heads)        # unet_pvc.py:98-99 (norm=False, flash=True)
        else:
            attention_fn = None
        with_se = pvd.get("use_se", True)
        sa_layers, sa_in_channels, c_sa, _ = create_sa_components(
            sa_blocks, extra_feature_channels=self.f_embed_dim, input_dim=self.input_dim, embed_dim=self.embed_dim,
            attention_fn=attention_fn, attention_layers=pvd.attentions, dropout=dropout, with_se=with_se,
            gn_groups=8, cond_dim=self.cond_emb_dim)
        self.sa_layers = nn.ModuleList(sa_layers)
        self.global_att = attention_fn(dim=c_sa) if attention_fn is not None else None
        sa_in_channels[0] = self.f_embed_dim + self.input_dim
        fp_layers, c_fp = create_fp_components(fp_blocks, in_channels=c_sa, sa_in_channels=sa_in_channels,
                                               embed_dim=self.embed_dim, dropout=dropout, with_se=with_se,
                                               gn_groups=8, cond_dim=self.cond_emb_dim)
        self.fp_layers = nn.ModuleList(fp_layers)
        layers, _ = create_mlp_components(c_fp, [pvd.get("out_mlp", 128), dropout, out_dim], classifier=True, dim=2)
        self.classifier = nn.ModuleList(layers)

    def get_timestep_embedding(self, timesteps, device):
        """unet_pvc.py:156-169 (frequency table built in float64 numpy, cast to fp32)."""
        if timesteps.dim() == 2 and timesteps.shape[1] == 1:
            timesteps = timesteps[:, 0]
        half = self.embed_dim // 2
        e = np.log(10000) / (half - 1)
        e = torch.from_numpy(np.exp(np.arange(0, half) * -e)).float().to(device)
        e = timesteps[:, None] * e[None, :]
        e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
        if self.embed_dim % 2 == 1:
            e = nn.functional.pad(e, (0, 1), "constant", 0)
        return e

    def forward(self, x, t, x_cond=None):
        if x_cond is not None:
            x = torch.cat([x, x_cond], dim=1)
        B, C, N = x.shape
        assert C == self.input_dim + self.extra_feature_channels, f"input dim: {C}, expected: {self.input_dim + self.extra_feature_channels}"
        coords = x[:, : self.input_dim].contiguous()
        feats = x[:, self.input_dim:].contiguous()
        if self.embed_feats is not None:
            feats = self.embed_feats(coords if self.extra_feature_channels == 0 else feats)
        data = PVCData(coords=coords, features=coords)
        if self.global_pnet is not None:
            data.cond = self.global_pnet(data)
        feats = torch.cat([coords, feats], dim=1)
        skips, coords_list = [feats], []
        temb = None
        if t is not None:
            if t.dim() == 0:
                t = t.view(1).expand(B)
            temb = self.embedf(self.get_timestep_embedding(t, x.device))[:, :, None].expand(-1, -1, N)
        data.features, data.time_emb = feats, temb
        for i, sa in enumerate(self.sa_layers):
            skips.append(data.features)
            coords_list.append(data.coords)
            if i > 0 and data.time_emb is not None:
                data.features = torch.cat([data.features, data.time_emb], dim=1)
            data = sa(data)
        skips.pop(1)
        if self.global_att is not None:
            data.features = self.global_att(data.features)
        for j, fp in enumerate(self.fp_layers):
            lower = torch.cat([data.features, data.time_emb], dim=1) if data.time_emb is not None else data.features
            data = fp(PVCData(features=skips[-1 - j], coords=coords_list[-1 - j], lower_coords=data.coords,
                              lower_features=lower, time_emb=data.time_emb, cond=data.cond))
        for l in self.classifier:
            if isinstance(l, SharedMLP):
                data.features = l(data).features
            else:
                data.features = l(data.features)
        return data.features
