"""Parameter containers + eager forward of the PVCNN building blocks, mirroring the reference's module tree so
that a reference checkpoint loads key-for-key (``models/pvcnn.py``, ``models/modules.py``).

Two execution paths share these modules:
  * ``forward`` here: channel-first eager path -- the point/voxel ops are this repo's CUDA kernels
    (``p2pb_b200.ops``), the dense contractions are torch library calls.  It is the reference-shaped path used to
    validate the fused engine on the GPU and as the op-level drop-in demonstration.
  * ``p2pb_b200.engine``: the fused channels-last engine (hand-written GEMM/conv kernels, CUDA graph) that
    ``P2PB.sample`` uses by default; it reads the weights out of these modules.
CUDA only: the ops raise on CPU tensors.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


@dataclass
class PVCData:
    """Same fields as the reference's ``PVCData`` (models/pvcnn.py:22-31)."""
    features: torch.Tensor
    coords: torch.Tensor = None
    cond_coords: torch.Tensor = None
    cond_features: torch.Tensor = None
    lower_coords: torch.Tensor = None
    lower_features: torch.Tensor = None
    time_emb: torch.Tensor = None
    cond: object = None


class Swish(nn.Module):
    def forward(self, x):  # modules.py:25-35
        return x * torch.sigmoid(x)


class AdaGN(nn.Module):
    """GroupNorm(x) * factor + bias, [factor, bias] = Linear(cond)  (modules.py:319-358)."""

    def __init__(self, num_channels: int, ctx_dim: int, ndim: int, num_groups: int = 8):
        super().__init__()
        self.ndim, self.n_channel = ndim, num_channels
        self.norm = nn.GroupNorm(num_groups, num_channels)
        self.emd = nn.Linear(ctx_dim, num_channels * 2)
        with torch.no_grad():  # variance-scaling fan_avg init, bias = (1, 0)  (modules.py:300-339)
            fan = (ctx_dim + 2 * num_channels) / 2.0
            bound = (3.0 / max(1.0, fan)) ** 0.5
            self.emd.weight.uniform_(-bound, bound)
            self.emd.bias[:num_channels] = 1
            self.emd.bias[num_channels:] = 0

    def forward(self, x, cond):
        e = self.emd(cond)
        e = e.view(e.shape[0], -1, *([1] * (x.dim() - 2)))
        factor, bias = e.chunk(2, 1)
        return self.norm(x) * factor + bias


class SE3d(nn.Module):
    def __init__(self, channel: int, reduction: int = 8):  # modules.py:362-378
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())
        self.channel = channel

    def forward(self, x):
        s = self.fc(x.mean(dim=(2, 3, 4)))
        return x * s.view(x.shape[0], x.shape[1], 1, 1, 1)


class LinearAttention(nn.Module):
    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32, verbose: bool = True):  # modules.py:165-194
        super().__init__()
        self.heads = heads
        hidden = dim_head * heads
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)

    def forward(self, x):
        B, C, N = x.shape
        qkv = self.to_qkv(x.unsqueeze(-1)).squeeze(-1)
        q, k, v = qkv.view(B, 3, self.heads, -1, N).unbind(1)
        k = k.softmax(dim=-1)
        ctx = torch.einsum("bhdn,bhen->bhde", k, v)
        out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, -1, N)
        return self.to_out(out.unsqueeze(-1)).squeeze(-1)


class Attention(nn.Module):
    """Softmax attention over the bottleneck tokens, the reference's ``attention_type: flash`` (modules.py:197-264 with norm=False,
    no time conditioning, no qk-norm, as unet_pvc.py:98-99 builds it; core = Attend, modules.py:77-162): parameters
    ``to_q [h*32, dim]``, ``to_kv [2*h*32, dim]``, ``to_out [dim, h*32]``, all bias-free Linears.  Takes / returns ``[B, C, N]``
    (the transposes of unet_pvc.py:239-241 are folded in)."""

    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32, **_unused):
        super().__init__()
        self.heads = heads
        hidden = dim_head * heads
        self.to_q = nn.Linear(dim, hidden, bias=False)
        self.to_kv = nn.Linear(dim, hidden * 2, bias=False)
        self.to_out = nn.Linear(hidden, dim, bias=False)

    def forward(self, x):
        B, C, N = x.shape
        t = x.transpose(1, 2)                                                   # [B, N, C] tokens
        q = self.to_q(t)
        k, v = self.to_kv(t).chunk(2, dim=-1)
        q, k, v = (z.view(B, N, self.heads, -1).transpose(1, 2) for z in (q, k, v))      # [B, h, N, d]
        attn = (torch.einsum("bhid,bhjd->bhij", q, k) * (q.shape[-1] ** -0.5)).softmax(dim=-1)
        out = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(1, 2).reshape(B, N, -1)
        return self.to_out(out).transpose(1, 2)


def _norm(num_channels: int, dim: int, gn_groups: int, cond_dim: int, affine: bool = True):
    if cond_dim > 0:
        return AdaGN(num_channels, cond_dim, dim, gn_groups)
    return nn.GroupNorm(gn_groups, num_channels, affine=affine)


class SharedMLP(nn.Module):
    """(1x1 conv -> norm -> Swish)*  (models/pvcnn.py:162-205); ``layers`` indices 0,1,2 / 3,4,5 ..."""

    def __init__(self, in_channels: int, out_channels, dim: int = 1, gn_groups: int = 8, cond_dim: int = 0,
                 affine: bool = True):
        super().__init__()
        conv = nn.Conv1d if dim == 1 else nn.Conv2d
        if not isinstance(out_channels, (list, tuple)):
            out_channels = [out_channels]
        layers: List[nn.Module] = []
        for oc in out_channels:
            layers += [conv(in_channels, oc, 1), _norm(oc, dim, gn_groups, cond_dim, affine), Swish()]
            in_channels = oc
        self.layers = nn.ModuleList(layers)

    def forward(self, data: PVCData) -> PVCData:
        x, cond = data.features, data.cond
        for l in self.layers:
            x = l(x, cond) if isinstance(l, AdaGN) and cond is not None else l(x)
        data.features = x
        return data


class Voxelization(nn.Module):
    """models/pvcnn.py:208-234, fused into one kernel (coordinate prep + CSR); returns the CSR too."""

    def __init__(self, resolution: int, normalize: bool = True, eps: float = 0):
        super().__init__()
        self.r, self.normalize, self.eps = int(resolution), normalize, eps

    def forward(self, features, coords):
        prep = ops.voxel_prep(coords.detach().contiguous(), self.r, self.normalize, self.eps)
        if features is None:
            return None, prep["norm_coords"]
        B, C, N = features.shape
        r = self.r
        # integer voxel coords back from the flat index (what the reference passes to avg_voxelize)
        ind = prep["ind"]
        vox = torch.stack([ind // (r * r), (ind // r) % r, ind % r], dim=1).int().contiguous()
        grid = ops.avg_voxelize(features.contiguous(), vox, r)[0]
        return grid.view(B, C, r, r, r), prep["norm_coords"]


class PVConv(nn.Module):
    """voxel branch (voxelize -> conv3d/AdaGN/Swish/conv3d/AdaGN/SE -> devoxelize) + point branch (pvcnn.py:237-334)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, resolution: int, normalize: bool = True,
                 eps: float = 0, with_se: bool = True, add_point_feat: bool = True, attention: bool = False,
                 attention_fn=None, dropout: float = 0.1, gn_groups: int = 8, cond_dim: int = 0, affine: bool = True):
        super().__init__()
        self.resolution = resolution
        self.voxelization = Voxelization(resolution, normalize=normalize, eps=eps)
        pad = kernel_size // 2
        vl: List[nn.Module] = [
            nn.Conv3d(in_channels, out_channels, kernel_size, stride=1, padding=pad),
            _norm(out_channels, 3, gn_groups, cond_dim, affine), Swish(), nn.Dropout(dropout),
            nn.Conv3d(out_channels, out_channels, kernel_size, stride=1, padding=pad),
            _norm(out_channels, 3, gn_groups, cond_dim, affine),
        ]
        if with_se:
            vl.append(SE3d(out_channels))
        self.voxel_layers = nn.ModuleList(vl)
        self.attn = attention_fn(out_channels) if attention else None
        self.add_point_feat = add_point_feat
        if add_point_feat:
            self.point_features = SharedMLP(in_channels, out_channels, gn_groups=gn_groups, cond_dim=cond_dim, affine=affine)

    def forward(self, data: PVCData) -> PVCData:
        coords, feats, cond = data.coords, data.features, data.cond
        v, norm_coords = self.voxelization(feats, coords)
        for l in self.voxel_layers:
            v = l(v, cond) if isinstance(l, AdaGN) else l(v)
        B, C = v.shape[:2]
        out = ops.trilinear_devoxelize(norm_coords, v.reshape(B, C, -1).contiguous(), self.resolution)
        if self.add_point_feat:
            out = out + self.point_features(data).features
        if self.attn is not None:
            out = self.attn(out)
        data.features = out
        return data


class PointNetSAModule(nn.Module):
    """FPS -> ball query -> grouping -> SharedMLP(dim=2) -> max over neighbours (pvcnn.py:337-424, 99-127)."""

    def __init__(self, num_centers: int, radius, num_neighbors, in_channels: int, out_channels,
                 include_coordinates: bool = True, gn_groups: int = 8, cond_dim: int = 0, affine_gn: bool = True):
        super().__init__()
        radius = radius if isinstance(radius, (list, tuple)) else [radius]
        num_neighbors = num_neighbors if isinstance(num_neighbors, (list, tuple)) else [num_neighbors] * len(radius)
        if not isinstance(out_channels, (list, tuple)):
            out_channels = [[out_channels]] * len(radius)
        elif not isinstance(out_channels[0], (list, tuple)):
            out_channels = [out_channels] * len(radius)
        self.radius, self.num_neighbors = list(radius), list(num_neighbors)
        self.include_coordinates = include_coordinates
        self.num_centers = num_centers
        self.mlps = nn.ModuleList([
            SharedMLP(in_channels + (3 if include_coordinates else 0), oc, dim=2, gn_groups=gn_groups, cond_dim=cond_dim,
                      affine=affine_gn) for oc in out_channels])
        self.out_channels = sum(oc[-1] for oc in out_channels)

    def forward(self, data: PVCData) -> PVCData:
        coords, feats, cond = data.coords[:, :3].contiguous(), data.features, data.cond
        _, centers = ops.furthest_point_sampling(coords, self.num_centers, return_centers=True)
        if data.time_emb is not None:
            data.time_emb = data.time_emb[:, :, : centers.shape[-1]]
        outs = []
        for radius, k, mlp in zip(self.radius, self.num_neighbors, self.mlps):
            nidx = ops.ball_query(centers, coords, radius, k)
            g = ops.grouping(coords, nidx) - centers.unsqueeze(-1)
            if feats is not None:
                g = torch.cat([g, ops.grouping(feats.contiguous(), nidx)], dim=1)
            outs.append(mlp(PVCData(features=g, cond=cond)).features.max(dim=-1).values)
        data.features = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
        data.coords = centers
        return data


class PointNetFPModule(nn.Module):
    """3-NN interpolation from the coarser level, skip concat, SharedMLP(dim=1) (pvcnn.py:427-467)."""

    def __init__(self, in_channels: int, out_channels, gn_groups: int = 8, cond_dim: int = 0, affine_gn: bool = True):
        super().__init__()
        self.mlp = SharedMLP(in_channels, out_channels, dim=1, gn_groups=gn_groups, cond_dim=cond_dim, affine=affine_gn)

    def forward(self, data: PVCData) -> PVCData:
        coords = data.coords[:, :3].contiguous()
        x, _, _ = ops.three_nn_interpolate(coords, data.lower_coords[:, :3].contiguous(), data.lower_features.contiguous())
        if data.features is not None:
            x = torch.cat([x, data.features], dim=1)
        if data.time_emb is not None:
            data.time_emb = data.time_emb[:, :, 0:1].expand(-1, -1, coords.shape[-1])
        data.features = x
        return self.mlp(data)


class MyGroupNorm(nn.Module):
    def __init__(self, num_groups: int, num_channels: int):  # pvcnn.py:745-763
        super().__init__()
        self.num_channels = num_channels - num_channels % num_groups
        self.num_groups = num_groups
        self.group_norm = nn.GroupNorm(num_groups, self.num_channels)

    def forward(self, x):
        if x.shape[1] == self.num_channels:
            return self.group_norm(x)
        return torch.cat([self.group_norm(x[:, : self.num_channels]), x[:, self.num_channels:]], dim=1)


class MLP(nn.Module):
    """conv(bias) -> MyGroupNorm(32) -> Swish  (pvcnn.py:766-823 with bn_first=False, activation=swish)."""

    def __init__(self, channels: Sequence[int], dim: int = 1, min_groups: int = 32):
        super().__init__()
        conv = nn.Conv1d if dim == 1 else nn.Conv2d
        layers: List[nn.Module] = []
        for i in range(1, len(channels)):
            layers += [conv(channels[i - 1], channels[i], kernel_size=1, bias=True), MyGroupNorm(min_groups, channels[i]), Swish()]
        self.mlp = nn.Sequential(*layers)

    def forward(self, data: PVCData) -> PVCData:
        data.features = self.mlp(data.features)
        return data


class ConditionedSharedMLPLayer(nn.Module):
    """pvcnn.py:826-902 as instantiated by Pnet2Stage (no time / cond embedding, no residual)."""

    def __init__(self, channels: Sequence[int], dim: int = 1):
        super().__init__()
        assert len(channels) > 2
        self.shared_mlp_0 = MLP([channels[0], channels[1]], dim=dim)
        self.shared_mlp_1 = MLP([channels[1], channels[2]], dim=dim)
        last, cin = [], channels[2]
        for oc in channels[3:]:
            last.append(MLP([cin, oc], dim=dim))
            cin = oc
        self.last_mlp_layers = nn.ModuleList(last)

    def forward(self, data: PVCData) -> PVCData:
        data = self.shared_mlp_1(self.shared_mlp_0(data))
        for l in self.last_mlp_layers:
            data = l(data)
        return data


class Pnet2Stage(nn.Module):
    """Global PointNet producing the conditioning vector (pvcnn.py:905-932)."""

    def __init__(self, mlp1: Sequence[int], mlp2: Sequence[int]):
        super().__init__()
        self.mlp1 = ConditionedSharedMLPLayer(list(mlp1), dim=2)
        self.mlp2 = ConditionedSharedMLPLayer([2 * mlp1[-1]] + list(mlp2), dim=2)

    def forward(self, x: PVCData):
        x.features = x.features.unsqueeze(-1)
        f = self.mlp1(x).features
        g = f.amax(dim=2, keepdim=True).expand(-1, -1, f.size(2), -1)
        x.features = torch.cat([f, g], dim=1)
        f = self.mlp2(x).features
        return f.amax(dim=2).squeeze(-1)


# ---------------------------------------------------------------------------------------------------------
# layer-spec builders (pvcnn.py:34-96, 478-741), reduced to what PVCNN2Unet instantiates
# ---------------------------------------------------------------------------------------------------------
def create_pvc_layer_params(npoints: int, channels, n_sa_blocks, n_fp_blocks, radius, voxel_resolutions,
                            downsample_factor: int = 4, centers=None):
    n = len(channels)
    sa_blocks = []
    for i in range(n - 1):
        nc = npoints // downsample_factor ** (i + 1) if centers is None else centers[i]
        if i != n - 2:
            sa_blocks.append([[channels[i], n_sa_blocks[i], voxel_resolutions[i]], [nc, radius[i], 32, [channels[i], channels[i + 1]]]])
        else:
            sa_blocks.append([None, [nc, radius[i], 32, [channels[i], channels[i], channels[i + 1]]]])
    fp_blocks = [
        [[channels[3], channels[3]], [channels[3], n_fp_blocks[3], voxel_resolutions[3]]],
        [[channels[3], channels[3]], [channels[3], n_fp_blocks[2], voxel_resolutions[2]]],
        [[channels[3], channels[2]], [channels[2], n_fp_blocks[1], voxel_resolutions[1]]],
        [[channels[2], channels[2], channels[1]], [channels[1], n_fp_blocks[0], voxel_resolutions[0]]],
    ]
    return sa_blocks, fp_blocks


def create_mlp_components(in_channels: int, out_channels, classifier: bool = False, dim: int = 2, gn_groups: int = 8,
                          cond_dim: int = 0):
    """Classifier head as built at unet_pvc.py:147-154 (dim=2 branch of pvcnn.py:478-525)."""
    assert dim == 2 and classifier
    layers: List[nn.Module] = []
    for oc in out_channels[:-1]:
        if oc < 1:
            layers.append(nn.Dropout(oc))
        else:
            layers.append(SharedMLP(in_channels, int(oc), gn_groups=gn_groups, cond_dim=cond_dim))
            in_channels = int(oc)
    layers.append(nn.Conv1d(in_channels, out_channels[-1], 1))
    return layers, out_channels[-1]


def create_sa_components(sa_blocks, extra_feature_channels: int, input_dim: int = 3, embed_dim: int = 64,
                         attention_fn=None, attention_layers=None, dropout: float = 0.1, with_se: bool = False,
                         gn_groups: int = 8, cond_dim: int = 0):
    in_channels = extra_feature_channels + input_dim
    sa_layers, sa_in_channels = [], []
    num_centers = None
    for c, (conv_cfg, sa_cfg) in enumerate(sa_blocks):
        k = 0
        sa_in_channels.append(in_channels)
        blocks: List[nn.Module] = []
        use_att = bool(attention_layers[c]) if attention_layers is not None else False
        if conv_cfg is not None:
            oc, num_blocks, res = conv_cfg
            for p in range(num_blocks):
                make = lambda ci: PVConv(ci, oc, kernel_size=3, resolution=int(res), attention=use_att and p == 0,
                                         attention_fn=attention_fn, dropout=dropout, with_se=with_se,
                                         gn_groups=gn_groups, cond_dim=cond_dim)
                if c == 0:          # only level 0 stacks n_sa_blocks PVConvs (pvcnn.py:615-618)
                    blocks.append(make(in_channels))
                elif k == 0:
                    blocks.append(make(in_channels + embed_dim))
                in_channels = oc
                k += 1
            extra_feature_channels = in_channels
        num_centers, radius, num_neighbors, ocs = sa_cfg
        blocks.append(PointNetSAModule(num_centers=num_centers, radius=radius, num_neighbors=num_neighbors,
                                       in_channels=extra_feature_channels + (embed_dim if k == 0 else 0),
                                       out_channels=[int(o) for o in ocs], include_coordinates=True,
                                       gn_groups=gn_groups, cond_dim=cond_dim))
        in_channels = extra_feature_channels = blocks[-1].out_channels
        sa_layers.append(blocks[0] if len(blocks) == 1 else nn.Sequential(*blocks))
    return sa_layers, sa_in_channels, in_channels, num_centers


def create_fp_components(fp_blocks, in_channels: int, sa_in_channels, embed_dim: int = 64, dropout: float = 0.1,
                         with_se: bool = False, gn_groups: int = 8, cond_dim: int = 0):
    fp_layers = []
    for j, (fp_cfg, conv_cfg) in enumerate(fp_blocks):
        ocs = tuple(int(o) for o in fp_cfg)
        blocks: List[nn.Module] = [PointNetFPModule(in_channels + sa_in_channels[-1 - j] + embed_dim, ocs,
                                                    gn_groups=gn_groups, cond_dim=cond_dim)]
        in_channels = ocs[-1]
        if conv_cfg is not None:
            oc, num_blocks, res = conv_cfg
            for _ in range(num_blocks):
                # no attention inside FP PVConvs: the reference's condition is never true (pvcnn.py:692,709)
                blocks.append(PVConv(in_channels, oc, kernel_size=3, resolution=int(res), attention=False,
                                     dropout=dropout, with_se=with_se, gn_groups=gn_groups, cond_dim=cond_dim))
                in_channels = oc
        fp_layers.append(blocks[0] if len(blocks) == 1 else nn.Sequential(*blocks))
    return fp_layers, in_channels
