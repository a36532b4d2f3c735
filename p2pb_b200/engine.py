"""Fused channels-last engine: one PVCNN U-Net evaluation + bridge update as ~250 hand-written kernels, all T
sampling steps captured in one CUDA graph.

Python here only *plans*: it packs the weights of a ``PVCNN2Unet`` once (K-major, channel-permuted, zero-padded to the
32-wide K chunks of the tcgen05 GEMM), allocates every intermediate buffer once (static shapes per (config, B, N)), and
enqueues kernels of ``libp2pb_b200.so`` through the C ABI on the current stream.  No torch op runs inside a network
evaluation.  Replaces the eager hot loop ``models/p2pb.py:215-262`` x ``models/unet_pvc.py:171-269``.

Layout: point features are rows ``[B*N, C]`` fp32 (C padded to a multiple of 32, feature channels FIRST, xyz after:
the reference concatenates ``[coords, features]``; weights are column-permuted at pack time), voxel grids are rows
``[B*r^3, C]``; coordinates stay ``[B,3,N]`` (what FPS / ball query / 3-NN read coalesced).

Fusions relative to the reference:
  * GroupNorm/AdaGN statistics come out of the GEMM/conv epilogue; normalise+AdaGN+Swish is one pass (or folded into the
    consumer: SE + AdaGN of the second voxel conv are applied inside the devoxelisation gather);
  * every ``torch.cat`` is a multi-segment GEMM operand; every ``cat[..., time_emb]`` in front of a 1x1 conv is a
    per-sample bias (``W[:, temb cols] @ temb``); in front of a voxel conv the time embedding is scatter-averaged by
    the voxelisation kernel itself;
  * geometry (FPS chain, ball queries, 3-NN, voxel CSRs) is computed once per evaluation and shared by all blocks.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import dense, ops
from ._lib import call, launch_count

_vp = ctypes.c_void_p
_f = ctypes.c_float


def _p(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _s():
    return _vp(torch.cuda.current_stream().cuda_stream)


def pad32(c: int) -> int:
    return (c + 31) // 32 * 32


def pad64(c: int) -> int:
    return (c + 63) // 64 * 64


class Options:
    """Build-time knobs of the engine.  There is ONE product path: these defaults.  The process environment does not change
    the arithmetic or the kernels that run; tests and the profiling tools under ``tools/`` flip a knob by assigning to
    ``engine.OPTIONS`` before an Engine is built (tests/test_engine_gpu.py shows each alternative is still correct).

    halo_f16 / gemm_f16: operands of the large-grid voxel convolutions (conv_halo) resp. of the r = 8 convs, the global
        PointNet and the shared-MLP chains are stored as IEEE half -- the same 10-bit mantissa a tf32 operand keeps inside
        the tensor core, twice the channels per byte and per MMA.  False keeps fp32 storage / kind::tf32 (also what a layer
        falls back to when its weights do not fit half's exponent range, see Engine._half_ok).
    group_project: first set-abstraction layer in gather-after-GEMM form;  pool_minmax: neighbourhood max-pool from the GEMM
        epilogue's column (max, min);  point_stream: PVConv point branch on a second stream;  chains: independent part-batch
        chains inside one graph (DualEngine);  no_graph: enqueue the kernels directly instead of replaying the captured
        CUDA graph (profilers that cannot see inside graphs)."""

    def __init__(self):
        self.halo_f16 = True
        self.gemm_f16 = True
        self.group_project = True
        self.pool_minmax = True
        self.point_stream = True
        self.chains = 1
        self.no_graph = False


OPTIONS = Options()


def halo_f16() -> bool:
    return OPTIONS.halo_f16


def gemm_f16() -> bool:
    return OPTIONS.gemm_f16


class _AdaGN:
    """Packed norm parameters: gamma/beta + offset of this layer's (factor, bias) block in the batched emd GEMM."""

    def __init__(self, gamma, beta, emd_off, groups):
        self.gamma, self.beta, self.emd_off, self.groups = gamma, beta, emd_off, groups


class Engine:
    # arithmetic of the contractions: 10-bit-mantissa operands (tf32 for the GEMMs / r=8 convs, IEEE half -- the same
    # mantissa -- for the r>=16 voxel convs), fp32 accumulation; everything else fp32
    dtype_name = "tf32"

    def __init__(self, p2pb, net, B: int, N: int, F: int, allow_half: bool = True):
        self.p2pb, self.net = p2pb, net
        self.B, self.N, self.F = B, N, F
        self.dev = next(net.parameters()).device
        self.E = net.embed_dim
        self.ind = net.input_dim
        assert self.ind == 3
        self.extra = net.extra_feature_channels
        assert F == self.extra, f"x_cond has {F} channels, model expects {self.extra}"
        # allow_half=False: this net produced values outside half's range in an earlier call (Engine._sample) -> fp32 / tf32 storage
        self.halo_f16, self.gemm_f16 = halo_f16() and allow_half, gemm_f16() and allow_half
        self.tf32_layers: List[str] = []      # layers whose WEIGHTS do not fit half's range and were packed as fp32 / tf32
        if self.halo_f16 or self.gemm_f16:
            self.dtype_name = "tf32+f16(10-bit mantissa operands), fp32 accumulate"
        self._folds: List[Tuple[str, torch.Tensor]] = []      # (bias2 buffer name, W[:, time columns]) of every temb fold
        self._emd_w: List[torch.Tensor] = []
        self._emd_b: List[torch.Tensor] = []
        self._emd_total = 0
        self._bufs: Dict[str, torch.Tensor] = {}
        self._graphs: Dict[tuple, tuple] = {}
        self.kernels_per_sample = 0
        with torch.no_grad():
            self._pack()

    # ------------------------------------------------------------------------------------------------ utils
    def zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=dtype, device=self.dev)

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.dev)

    def _w(self, t):
        return t.detach().to(self.dev, torch.float32).contiguous()

    def _half_ok(self, name: str, *ws) -> bool:
        """IEEE half keeps tf32's 10 mantissa bits but only 5 exponent bits: a weight tensor is stored as half only if its largest
        magnitude is finite in half (<= 6e4) and well inside the normal range (>= 1e-3: below, a growing share of the values is
        sub-normal in half and loses relative precision).  Otherwise this layer keeps fp32 storage / kind::tf32 -- the reference's
        arithmetic -- and is listed in ``tf32_layers``."""
        for w in ws:
            m = float(w.detach().abs().max())
            if not (1e-3 <= m <= 6.0e4):
                self.tf32_layers.append(f"{name} (max|w| = {m:.3g})")
                return False
        return True

    def _norm(self, mod, groups=8) -> _AdaGN:
        """AdaGN (norm + emd) or plain GroupNorm / MyGroupNorm."""
        if hasattr(mod, "emd"):
            off = self._emd_total
            self._emd_w.append(self._w(mod.emd.weight))
            self._emd_b.append(self._w(mod.emd.bias))
            self._emd_total += mod.emd.weight.shape[0]
            return _AdaGN(self._w(mod.norm.weight), self._w(mod.norm.bias), off, mod.norm.num_groups)
        gn = mod.group_norm if hasattr(mod, "group_norm") else mod
        return _AdaGN(self._w(gn.weight), self._w(gn.bias), -1, gn.num_groups)

    def _pack_rows_w(self, w, col_map, k_pad) -> torch.Tensor:
        """[O, I(,1..)] -> [O, k_pad] with source column blocks moved: col_map = [(src0, n, dst0), ...]."""
        w = self._w(w).reshape(w.shape[0], w.shape[1])
        out = self.zeros(w.shape[0], k_pad)
        for src, n, dst in col_map:
            out[:, dst:dst + n] = w[:, src:src + n]
        return out.contiguous()

    # ------------------------------------------------------------------------------------------------ packing
    def _pack_pvconv(self, mod, c_in: int, temb_in: bool, coords_first: bool, name: str = "pvconv", ename: str = "pv"):
        """c_in = valid channels of the incoming rows (features[, xyz]); temb_in: 64 time channels follow in the
        reference's channel order.  coords_first: reference input order is [xyz, feats] (level 0) -> ours [feats, xyz]."""
        E = self.E if temb_in else 0
        if getattr(mod, "attn", None) is not None:
            raise NotImplementedError("PVConv with attention=True: no shipped config builds one (pvcnn.py:692,709 shadow "
                                      "`fp_blocks`), the fused engine does not implement it")
        conv1, n1, conv2, n2 = mod.voxel_layers[0], mod.voxel_layers[1], mod.voxel_layers[4], mod.voxel_layers[5]
        se = mod.voxel_layers[6] if len(mod.voxel_layers) > 6 else None
        cout = conv1.out_channels
        cin_ref = conv1.in_channels
        assert cin_ref == c_in + E
        if coords_first:
            perm = list(range(3, c_in)) + [0, 1, 2]
        else:
            perm = list(range(c_in))
        perm_full = perm + list(range(c_in, c_in + E))
        cp = pad32(c_in + E)
        P = {"cout": cout, "cin": c_in, "E": E, "cp": cp, "r": int(mod.resolution)}
        halo = int(mod.resolution) >= 16 and cout <= 128 and cout % 32 == 0   # large grid / few channels: conv_halo.cu
        P["halo"] = halo
        w_ok = (self.halo_f16 or self.gemm_f16) and self._half_ok(name + ".voxel_layers", conv1.weight, conv2.weight)
        if halo and self.halo_f16 and w_ok:
            P["cp"] = cp = pad64(c_in + E)
            P["w1"] = dense.pack_conv3d_weight(self._w(conv1.weight), cp, perm_full).half()
            P["w2"] = dense.pack_conv3d_weight(self._w(conv2.weight), pad64(cout)).half()
        elif not halo and self.gemm_f16 and cout % 32 == 0 and w_ok:
            P["cp"] = cp = pad64(c_in + E)
            P["w1"] = dense.pack_conv3d_weight(self._w(conv1.weight), cp, perm_full).half()
            P["w2"] = dense.pack_conv3d_weight(self._w(conv2.weight), pad64(cout)).half()
        else:
            P["w1"] = dense.pack_conv3d_weight(self._w(conv1.weight), cp, perm_full)
            P["w2"] = dense.pack_conv3d_weight(self._w(conv2.weight), pad32(cout))
        P["b1"] = self._w(conv1.bias)
        P["n1"] = self._norm(n1)
        P["b2"] = self._w(conv2.bias)
        P["n2"] = self._norm(n2)
        if se is not None:
            P["se0"], P["se2"] = self._w(se.fc[0].weight), self._w(se.fc[2].weight)
        pf = mod.point_features.layers
        wp = self._w(pf[0].weight).reshape(cout, cin_ref)
        kp = pad32(c_in)
        P["wp"] = self.zeros(cout, kp)
        P["wp"][:, :c_in] = wp[:, perm]
        P["wp"] = P["wp"].contiguous()
        P["bp"] = self._w(pf[0].bias)
        if E:       # cat[features, time_emb] in front of the point branch's 1x1 conv == per-sample bias (p2pb_step_vectors)
            self._folds.append((f"{ename}.ptb", wp[:, c_in:c_in + E].contiguous()))
        P["np"] = self._norm(pf[1])
        return P

    def _pack_mlp(self, layers, first_cols, k_pad_first, temb_cols=None, half_first=False, name: str = "mlp", ename: str = "mlp"):
        """SharedMLP -> list of dicts; first layer's columns remapped by first_cols; optional temb fold block.
        With gemm_f16 the layers after the first take IEEE-half operands (their input is a GroupNorm+Swish output written
        by this engine); the first layer too when its input rows are produced as half (half_first: grouped rows)."""
        out = []
        i = 0
        while i < len(layers):
            conv, nm = layers[i], layers[i + 1]
            o, c = conv.weight.shape[:2]
            L = {"cout": o}
            if i == 0:
                L["w"] = self._pack_rows_w(conv.weight, first_cols, k_pad_first)
                if half_first and self.gemm_f16:
                    L["w"] = L["w"].half()
                if temb_cols is not None:
                    w = self._w(conv.weight).reshape(o, c)
                    L["tb_name"] = f"{ename}.0.tb"
                    self._folds.append((L["tb_name"], w[:, temb_cols[0]:temb_cols[0] + temb_cols[1]].contiguous()))
            elif self.gemm_f16 and o % 32 == 0 and self._half_ok(f"{name}.layers.{i}", conv.weight):
                L["w"] = self._pack_rows_w(conv.weight, [(0, c, 0)], pad64(c)).half()
            else:
                L["w"] = self._pack_rows_w(conv.weight, [(0, c, 0)], pad32(c))
            L["b"] = self._w(conv.bias)
            L["n"] = self._norm(nm)
            out.append(L)
            i += 3
        return out

    def _pack(self):
        net, E = self.net, self.E
        self.fe = net.f_embed_dim
        fe = self.fe
        assert fe % 4 == 0
        W = {}
        # time embedding MLP
        W["tw0"], W["tb0"] = self._w(net.embedf[0].weight), self._w(net.embedf[0].bias)
        W["tw2"], W["tb2"] = self._w(net.embedf[2].weight), self._w(net.embedf[2].bias)
        # feature embedding
        if net.embed_feats is not None:
            ef = net.embed_feats
            cin = ef[0].weight.shape[1]
            W["ef0"] = self._pack_rows_w(ef[0].weight, [(0, cin, 0)], pad32(cin))
            W["ef0b"] = self._w(ef[0].bias)
            W["efn"] = self._norm(ef[1])
            W["ef3"] = self._pack_rows_w(ef[3].weight, [(0, fe, 0)], pad32(fe))
            W["ef3b"] = self._w(ef[3].bias)
        # global PointNet
        self.cond_dim = net.cond_emb_dim
        if net.global_pnet is not None:
            gp = net.global_pnet
            m = [gp.mlp1.shared_mlp_0.mlp, gp.mlp1.shared_mlp_1.mlp, gp.mlp2.shared_mlp_0.mlp, gp.mlp2.shared_mlp_1.mlp]
            G = []
            for j, seq in enumerate(m):
                conv, gn = seq[0], seq[1]
                o, c = conv.weight.shape[:2]
                L = {"cout": o, "b": self._w(conv.bias), "n": self._norm(gn)}
                if j == 0:      # one storage type for the whole chain; (the column max/min epilogue needs whole 128-row tiles per sample)
                    pnet_f16 = self.gemm_f16 and self.N % 128 == 0 and self._half_ok("global_pnet", *[q[0].weight for q in m])
                padk = pad64 if pnet_f16 else pad32
                if j == 2:  # input = cat[point feature (c/2), global max (c/2)]: second half becomes a per-sample bias
                    h = c // 2
                    L["w"] = self._pack_rows_w(conv.weight, [(0, h, 0)], padk(h))
                    L["w_g"] = self._w(conv.weight).reshape(o, c)[:, h:].contiguous()
                else:
                    L["w"] = self._pack_rows_w(conv.weight, [(0, c, 0)], padk(c))
                if pnet_f16:
                    L["w"] = L["w"].half()
                G.append(L)
            W["pnet"] = G
        # SA levels
        sa = []
        c_feat = fe + 3
        self.Ns = [self.N]
        n_levels = len(net.sa_layers)
        for i, blk in enumerate(net.sa_layers):
            mods = list(blk) if isinstance(blk, torch.nn.Sequential) else [blk]
            pvs, sam = mods[:-1], mods[-1]
            L = {"pv": [], "skip_c": c_feat}
            c_cur = c_feat
            for k, pv in enumerate(pvs):
                L["pv"].append(self._pack_pvconv(pv, c_cur, temb_in=(i > 0 and k == 0), coords_first=(i == 0 and k == 0),
                                                 name=f"sa_layers.{i}.{k}", ename=f"sa{i}.pv{k}"))
                c_cur = L["pv"][-1]["cout"]
            temb_sa = (len(pvs) == 0 and i > 0)
            # SA-module MLP: reference grouped input = [rel xyz (3), features (c_cur) (+ temb 64)]; ours [features, rel xyz]
            if i == 0 and len(pvs) == 0:
                raise NotImplementedError("level 0 without PVConv")
            assert c_cur % 4 == 0
            sa_half = (self.gemm_f16 and sam.mlps[0].layers[0].weight.shape[0] % 32 == 0
                       and self._half_ok(f"sa_layers.{i}.mlps.0.layers.0", sam.mlps[0].layers[0].weight))
            L["mlp"] = self._pack_mlp(sam.mlps[0].layers, [(3, c_cur, 0), (0, 3, c_cur)],
                                      pad64(c_cur + 3) if sa_half else pad32(c_cur + 3),
                                      temb_cols=(3 + c_cur, E) if temb_sa else None, half_first=sa_half, name=f"sa_layers.{i}.mlps.0",
                                      ename=f"sa{i}.mlp")
            # gather-after-GEMM form of the first layer (p2pb_group_project): feature part per point, coordinate part in fp32
            conv0 = sam.mlps[0].layers[0]
            o0 = conv0.weight.shape[0]
            w0 = self._w(conv0.weight).reshape(o0, -1)
            L["proj"] = (self.gemm_f16 and o0 % 32 == 0 and int(sam.num_neighbors[0]) == 32 and len(L["mlp"]) > 1
                         and L["mlp"][1]["w"].dtype == torch.float16 and OPTIONS.group_project)
            if L["proj"]:
                L["w_f"] = self.zeros(o0, pad32(c_cur))
                L["w_f"][:, :c_cur] = w0[:, 3:3 + c_cur]
                L["w_f"] = L["w_f"].contiguous()
                L["w_x"] = w0[:, 0:3].contiguous()
            L["centers"], L["radius"], L["K"] = sam.num_centers, float(sam.radius[0]), int(sam.num_neighbors[0])
            L["c_grp"] = c_cur
            sa.append(L)
            c_feat = L["mlp"][-1]["cout"]
            self.Ns.append(sam.num_centers)
        W["sa"] = sa
        # bottleneck attention
        if net.global_att is not None:
            ga = net.global_att
            if hasattr(ga, "to_qkv"):          # LinearAttention (modules.py:165-194)
                W["att_kind"] = "linear"
                W["att_qkv"] = self._pack_rows_w(ga.to_qkv.weight, [(0, c_feat, 0)], pad32(c_feat))
                W["att_outb"] = self._w(ga.to_out.bias)
            else:                              # Attention / Attend (modules.py:197-264, 77-162): [to_q | to_kv] as one GEMM, no biases
                W["att_kind"] = "softmax"
                W["att_qkv"] = self._pack_rows_w(torch.cat([ga.to_q.weight, ga.to_kv.weight], 0), [(0, c_feat, 0)], pad32(c_feat))
                W["att_outb"] = None
            hid = ga.to_out.weight.shape[1]
            W["att_out"] = self._pack_rows_w(ga.to_out.weight, [(0, hid, 0)], pad32(hid))
            W["heads"] = ga.heads
        # FP levels
        fp = []
        c_low = c_feat
        for j, blk in enumerate(net.fp_layers):
            mods = list(blk) if isinstance(blk, torch.nn.Sequential) else [blk]
            fpm, pvs = mods[0], mods[1:]
            lvl = n_levels - 1 - j
            c_skip = sa[lvl]["skip_c"]
            kp_skip = pad32(c_skip)
            # reference input = [interp(c_low + temb E), skip(c_skip)]; skip at level 0 is [xyz(3), feats(fe)] -> ours [feats, xyz]
            if lvl == 0:
                skip_cols = [(c_low + E + 3, c_skip - 3, c_low), (c_low + E, 3, c_low + c_skip - 3)]
            else:
                skip_cols = [(c_low + E, c_skip, c_low)]
            assert c_low % 32 == 0
            L = {"mlp": self._pack_mlp(fpm.mlp.layers, [(0, c_low, 0)] + skip_cols, c_low + kp_skip, temb_cols=(c_low, E),
                                       name=f"fp_layers.{j}.mlp", ename=f"fp{j}.mlp"),
                 "c_low": c_low, "kp_skip": kp_skip, "lvl": lvl, "pv": []}
            c_cur = L["mlp"][-1]["cout"]
            for pv in pvs:
                L["pv"].append(self._pack_pvconv(pv, c_cur, temb_in=False, coords_first=False, name=f"fp_layers.{j}.{1 + len(L['pv'])}",
                                                 ename=f"fp{j}.pv{len(L['pv'])}"))
                c_cur = L["pv"][-1]["cout"]
            fp.append(L)
            c_low = c_cur
        W["fp"] = fp
        # classifier
        cl = net.classifier
        W["cls"] = self._pack_mlp(cl[0].layers, [(0, c_low, 0)], pad32(c_low), name="classifier.0")
        om = cl[-1].weight.shape[1]
        assert cl[-1].weight.shape[0] == 3 and om % 4 == 0
        W["cls_out"] = self._w(cl[-1].weight).reshape(3, om).contiguous()       # fp32 [3, C]: p2pb_head_bridge
        W["cls_outb"] = self._w(cl[-1].bias)
        # batched AdaGN emd weights
        if self._emd_total:
            W["emd_w"] = torch.cat(self._emd_w, 0).contiguous()
            W["emd_b"] = torch.cat(self._emd_b, 0).contiguous()
            pad = (-self._emd_total) % 128
            if pad:
                W["emd_w"] = torch.cat([W["emd_w"], self.zeros(pad, W["emd_w"].shape[1])], 0).contiguous()
                W["emd_b"] = torch.cat([W["emd_b"], self.zeros(pad)], 0).contiguous()
            self._emd_ld = W["emd_w"].shape[0]
        self.W = W
        # temb folds: one dense [B, cout] bias2 buffer each, all produced by ONE p2pb_step_vectors launch per step
        self._fold_R = sum(w.shape[0] for _, w in self._folds)
        if self._folds:
            W["fold_w"] = torch.cat([w for _, w in self._folds], 0).contiguous()
            ptr, stride = [], []
            for nm, w in self._folds:
                b = self.buf(nm, self.B, w.shape[0])
                ptr += [b.data_ptr() + 4 * o for o in range(w.shape[0])]
                stride += [w.shape[0]] * w.shape[0]
            W["fold_ptr"] = torch.tensor(ptr, dtype=torch.int64, device=self.dev)
            W["fold_stride"] = torch.tensor(stride, dtype=torch.int32, device=self.dev)

    # ------------------------------------------------------------------------------------------------ buffers
    def buf(self, name: str, *shape, dtype=torch.float32) -> torch.Tensor:
        """Named, zero-initialised, allocated once (padding columns stay zero forever)."""
        t = self._bufs.get(name)
        if t is None:
            t = self.zeros(*shape, dtype=dtype)
            self._bufs[name] = t
        assert tuple(t.shape) == tuple(shape), (name, t.shape, shape)
        return t

    def padded(self, name: str, B: int, C: int, r: int, dtype=torch.float32) -> torch.Tensor:
        """Zero-bordered padded-linear conv input [B*(r+2)^3 + slack, C] (conv_halo.cu); borders stay zero forever."""
        t = self._bufs.get(name)
        if t is None:
            t = dense.alloc_padded(B, C, r, self.dev, dtype)
            self._bufs[name] = t
        return t

    # ------------------------------------------------------------------------------------------------ primitives
    def gemm(self, name, segs, ks, w, bias, n_out, rows_per_sample, bias2=None, want_stats=True, out=None, minmax=False,
             store=True):
        """rows GEMM + (sum, sum^2) statistics in the format gn_coef expects -> (raw, stats, tiles).
        minmax: also per-tile column (max, min) -> self.last_colmm; store=False: statistics only (raw is None)."""
        M = segs[0].shape[0]
        raw = None
        if store:
            raw = out if out is not None else self.buf(name + ".raw", M, n_out)
        stats, tiles = None, 0
        fused = want_stats and rows_per_sample % 128 == 0
        if want_stats:
            tiles = rows_per_sample // 32 if fused else 1      # the GEMM epilogue writes one partial per 32-row block
            stats = self.buf(name + ".stats", (M // rows_per_sample) * tiles, n_out, 2)
        colmm = None
        if minmax:
            assert fused and n_out % 32 == 0
            colmm = self.buf(name + ".colmm", (M // rows_per_sample) * tiles, n_out, 2)
        self.last_colmm = colmm
        dense.gemm_rows(segs, w, bias, bias2, rows_per_sample if bias2 is not None else 0, out=raw,
                        stats=stats if fused else None, ks=ks, colmm=colmm, store=store)
        if want_stats and not fused:
            call("p2pb_col_stats", _p(raw), int(raw.stride(0)), M // rows_per_sample, rows_per_sample, n_out, _p(stats), _s())
        return raw, stats, tiles

    def coef(self, name, stats, tiles, nrm: _AdaGN, C, rows_per_sample, want_mean=False):
        B = self.B
        A, Bc = self.buf(name + ".A", B, C), self.buf(name + ".B", B, C)
        ym = self.buf(name + ".ym", B, C) if want_mean else None
        emd = self.emd_all if nrm.emd_off >= 0 else None
        call("p2pb_gn_coef", _p(stats), tiles, B, C, nrm.groups, rows_per_sample, _p(nrm.gamma), _p(nrm.beta), _p(emd),
             self._emd_ld if emd is not None else 0, max(nrm.emd_off, 0), _f(1e-5), _p(A), _p(Bc), _p(ym), _s())
        return A, Bc, ym

    def act(self, x, A, Bc, rows_per_sample, C, out, act=1, pool=1, gmax=None):
        if out is not None and out.dtype == torch.float16:      # operand of a half GEMM / conv
            assert pool == 1 and gmax is None
            call("p2pb_affine_act_f16", _p(x), int(x.stride(0)), _p(A), _p(Bc), rows_per_sample, x.shape[0], C, act, _p(out),
                 int(out.stride(0)), _s())
            return
        call("p2pb_affine_act", _p(x), int(x.stride(0)), _p(A), _p(Bc), rows_per_sample, x.shape[0], C, act, pool,
             _p(out), int(out.stride(0)) if out is not None else 0, _p(gmax), _s())

    def linear(self, x, w, bias, act, out, K=None, O=None):
        call("p2pb_linear_small", _p(x), int(x.stride(0)), _p(w), int(w.stride(0)), _p(bias), x.shape[0], K or w.shape[1],
             O or w.shape[0], act, _p(out), int(out.stride(0)), _s())

    def mlp_chain(self, name, layers, segs, ks, rows_per_sample, temb, final_pool=1, final_out=None, li0=0):
        """(GEMM -> GN/AdaGN coef -> Swish)* ; the last layer's activation optionally max-pools `final_pool` rows."""
        B = self.B
        x_segs, x_ks = segs, ks
        for li, L in enumerate(layers):
            nm = f"{name}.{li + li0}"
            bias2 = self.buf(L["tb_name"], B, L["cout"]) if (li == 0 and "tb_name" in L) else None     # filled by step_vectors
            last = li == len(layers) - 1
            # neighbourhood max-pool over K = 32 grouped rows == one 32-row block of the GEMM epilogue's column (max, min):
            # the last layer's [B*M*32, C] output is never written and the pooling pass disappears (p2pb_pool32_minmax)
            pool_mm = (last and final_pool == 32 and final_out is None and rows_per_sample % 128 == 0 and L["cout"] % 32 == 0
                       and OPTIONS.pool_minmax)
            raw, stats, tiles = self.gemm(nm, x_segs, x_ks, L["w"], L["b"], L["cout"], rows_per_sample, bias2=bias2,
                                          minmax=pool_mm, store=not pool_mm)
            colmm = self.last_colmm
            A, Bc, _ = self.coef(nm, stats, tiles, L["n"], L["cout"], rows_per_sample)
            if pool_mm:
                M_rows = (x_segs[0].shape[0]) // 32
                out = self.buf(nm + ".pool", M_rows, L["cout"])
                call("p2pb_pool32_minmax", _p(colmm), B, M_rows // B, L["cout"], _p(A), _p(Bc), 1, _p(out), int(out.stride(0)), _s())
                return out
            if last and final_out is not None:
                out = final_out
            elif last and final_pool > 1:
                out = self.buf(nm + ".pool", raw.shape[0] // final_pool, L["cout"])
            elif not last and layers[li + 1]["w"].dtype == torch.float16:      # operand of a half GEMM
                out = self.buf(nm + ".act", raw.shape[0], pad64(L["cout"]), dtype=torch.float16)
            else:
                out = self.buf(nm + ".act", raw.shape[0], pad32(L["cout"]))
            self.act(raw, A, Bc, rows_per_sample, L["cout"], out, act=1, pool=final_pool if last else 1)
            x_segs, x_ks = [out], [out.shape[1]] if not last else None
        return out

    def sa_projected(self, name, L, feats, coords_pts, coords_ctr, nidx, temb, n_pts, M, K, cg):
        """Set-abstraction shared MLP with the first layer in gather-after-GEMM form (no grouped tensor, no first-layer output):
        Pf = features @ Wf^T + bias (+ temb fold) per point; statistics and activation of v = Pf[idx] + Wx.(xyz[idx]-centre)
        are two passes of p2pb_group_project; the remaining layers run as usual on the half activation rows."""
        B = self.B
        L0 = L["mlp"][0]
        c0 = L0["cout"]
        nm = f"{name}.mlp.0"
        bias2 = self.buf(L0["tb_name"], B, c0) if "tb_name" in L0 else None       # filled by step_vectors
        pf, _, _ = self.gemm(nm + ".pf", [feats], [pad32(cg)], L["w_f"], L0["b"], c0, n_pts, bias2=bias2, want_stats=False)
        stats = self.buf(nm + ".stats", B * M, c0, 2)
        args = (_p(pf), int(pf.stride(0)), _p(L["w_x"]), _p(coords_pts), _p(coords_ctr), _p(nidx))
        call("p2pb_group_project", *args, _vp(0), _vp(0), _p(stats), _vp(0), 0, B, c0, n_pts, M, K, 0, _s())
        A, Bc, _ = self.coef(nm, stats, M, L0["n"], c0, M * K)
        act0 = self.buf(nm + ".act", B * M * K, pad64(c0), dtype=torch.float16)
        call("p2pb_group_project", *args, _p(A), _p(Bc), _vp(0), _p(act0), int(act0.stride(0)), B, c0, n_pts, M, K, 1, _s())
        return self.mlp_chain(f"{name}.mlp", L["mlp"][1:], [act0], [act0.shape[1]], M * K, temb, final_pool=K, li0=1)

    def pvconv(self, name, P, feats, coords_lvl, prep, temb, n_pts):
        """PVConv (pvcnn.py:306-334): voxel branch + point branch -> rows [B*n_pts, cout]."""
        B, r, cout, cp = self.B, P["r"], P["cout"], P["cp"]
        r3 = r ** 3
        halo = P["halo"]
        # point branch (1x1 conv + AdaGN coefficients): independent of the voxel branch until the devoxelisation, so it runs
        # on a second stream and fills the gaps the voxel branch leaves (conv tails, the HBM-bound activation pass)
        main = torch.cuda.current_stream()
        pt_done = None
        pstream = self._point_stream() if OPTIONS.point_stream else None
        if pstream is not None:
            fork = torch.cuda.Event()
            fork.record(main)
            pstream.wait_event(fork)
        with torch.cuda.stream(pstream if pstream is not None else main):
            bias2 = self.buf(f"{name}.ptb", B, cout) if P["E"] else None       # filled by step_vectors
            kp = P["wp"].shape[1]
            praw, pst, ptiles = self.gemm(f"{name}.pt", [feats], [kp], P["wp"], P["bp"], cout, n_pts, bias2=bias2)
            pA, pB, _ = self.coef(f"{name}.np", pst, ptiles, P["np"], cout, n_pts)
            if pstream is not None:
                pt_done = torch.cuda.Event()
                pt_done.record(pstream)
        raw1 = self.buf(f"{name}.raw1", B * r3, cout)
        raw2 = self.buf(f"{name}.raw2", B * r3, cout)
        tvox = _p(temb if P["E"] else None)
        if halo:
            f16 = P["w1"].dtype == torch.float16
            dt = torch.float16 if f16 else torch.float32
            sfx = "_f16" if f16 else ""
            c2 = pad64(cout) if f16 else pad32(cout)
            _, _, tiles1 = dense.halo_layout(r, cout, f16, cin=cp)          # conv1: padded Cin = cp, conv2: Cin = c2
            _, _, tiles = dense.halo_layout(r, cout, f16, cin=c2)
            grid = self.padded(f"{name}.grid", B, cp, r, dt)
            # the grid stays all-zero between evaluations: write the occupied voxels, convolve, zero them again
            vox_args = (_p(feats), int(feats.stride(0)), P["cin"], tvox, P["E"], _p(prep["order"]), _p(prep["ind"]),
                        _p(prep["start"]), _p(prep["cnt"]), _p(grid), cp, B, n_pts, r)
            call("p2pb_voxelize_padded_sparse" + sfx, *vox_args, 0, _s())
            st1 = self.buf(f"{name}.st1", B * tiles1, cout, 2)
            dense.conv3d_halo(grid, P["w1"], P["b1"], B, r, cp, cout, out=raw1, stats=st1, cin_valid=P["cin"] + P["E"])
            call("p2pb_voxelize_padded_sparse" + sfx, *vox_args, 1, _s())
            A1, B1, _ = self.coef(f"{name}.n1", st1, tiles1, P["n1"], cout, r3)
            act1 = self.padded(f"{name}.act1", B, c2, r, dt)
            if f16:
                call("p2pb_affine_act_padded_f16", _p(raw1), cout, _p(A1), _p(B1), B, cout, r, _p(act1), c2, _s())
            else:
                call("p2pb_affine_act_padded", _p(raw1), cout, _p(A1), _p(B1), B, cout, r, _p(act1), _s())
            st2 = self.buf(f"{name}.st2", B * tiles, cout, 2)
            dense.conv3d_halo(act1, P["w2"], P["b2"], B, r, c2, cout, out=raw2, stats=st2, cin_valid=cout)
        else:
            tiles = r3 // 32
            f16 = P["w1"].dtype == torch.float16
            dt = torch.float16 if f16 else torch.float32
            c2 = pad64(cout) if f16 else pad32(cout)
            grid = self.buf(f"{name}.grid", B * r3, cp, dtype=dt)
            call("p2pb_voxelize_cl_f16" if f16 else "p2pb_voxelize_cl", _p(feats), int(feats.stride(0)), P["cin"], tvox, P["E"],
                 _p(prep["order"]), _p(prep["start"]), _p(prep["cnt"]), _p(grid), cp, B, n_pts, r, _s())
            st1 = self.buf(f"{name}.st1", B * tiles, cout, 2)
            dense.conv3d_cl(grid, P["w1"], P["b1"], B, r, cp, cout, out=raw1, stats=st1)
            A1, B1, _ = self.coef(f"{name}.n1", st1, tiles, P["n1"], cout, r3)
            act1 = self.buf(f"{name}.act1", B * r3, c2, dtype=dt)
            self.act(raw1, A1, B1, r3, cout, act1, act=1)
            st2 = self.buf(f"{name}.st2", B * tiles, cout, 2)
            dense.conv3d_cl(act1, P["w2"], P["b2"], B, r, c2, cout, out=raw2, stats=st2)
        A2, B2, ym = self.coef(f"{name}.n2", st2, tiles, P["n2"], cout, r3, want_mean="se0" in P)
        se = None
        if "se0" in P:
            se = self.buf(f"{name}.se", B, cout)
            call("p2pb_se_excite", _p(ym), _p(P["se0"]), _p(P["se2"]), B, cout, P["se0"].shape[0], _p(se), _s())
        if pt_done is not None:
            torch.cuda.current_stream().wait_event(pt_done)
        out = self.buf(f"{name}.out", B * n_pts, cout)
        call("p2pb_devox_cl", _p(prep["norm_coords"]), _p(raw2), cout, _p(A2), _p(B2), _p(se), _p(praw), int(praw.stride(0)),
             _p(pA), _p(pB), _p(out), cout, B, cout, n_pts, r, _s())
        return out

    def voxel_prep(self, cache, lvl, coords, r):
        key = (lvl, r)
        if key not in cache:
            B, _, n = coords.shape
            nc = self.buf(f"prep{key}.nc", B, 3, n)
            ind = self.buf(f"prep{key}.ind", B, n, dtype=torch.int32)
            order = self.buf(f"prep{key}.order", B, n, dtype=torch.int32)
            start = self.buf(f"prep{key}.start", B, r ** 3, dtype=torch.int32)
            cnt = self.buf(f"prep{key}.cnt", B, r ** 3, dtype=torch.int32)
            call("p2pb_voxel_prep", _p(coords), B, n, r, 1, _f(0.0), _p(nc), _p(ind), _p(order), _p(start), _p(cnt), _s())
            cache[key] = {"norm_coords": nc, "order": order, "start": start, "cnt": cnt, "ind": ind}
        return cache[key]

    def _point_stream(self):
        if getattr(self, "_pstream", None) is None:
            self._pstream = torch.cuda.Stream(device=self.dev)
        return self._pstream

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.dev)
        return self._side

    def _geometry(self, xt):
        """FPS chain, ball-query indices, 3-NN indices/weights and the voxel CSR of every (level, resolution).
        The level-0 voxel CSRs depend on the input coordinates only and come first: an event after them lets the main
        stream start the first PVConv while the (latency-bound, one CTA per patch) FPS chain is still running."""
        B, N, W = self.B, self.N, self.W
        Ns = self.Ns
        n_levels = len(W["sa"])
        coords = [xt]
        preps: Dict[tuple, dict] = {}
        for P in W["sa"][0]["pv"]:
            self.voxel_prep(preps, 0, xt, P["r"])
        for L in W["fp"]:
            if L["lvl"] == 0:
                for P in L["pv"]:
                    self.voxel_prep(preps, 0, xt, P["r"])
        ev_prep0 = torch.cuda.Event()
        ev_prep0.record(torch.cuda.current_stream())
        # per level: FPS -> ball query -> event, so that set-abstraction level i only waits for ITS centres / neighbour lists
        # (at N = 8192 the level-0 FPS + ball query take 2.3 ms, the deeper levels another 0.4 ms)
        nidx, ev_level = [], []
        for i, L in enumerate(W["sa"]):
            M = Ns[i + 1]
            idx = self.buf(f"fps{i}.idx", B, M, dtype=torch.int32)
            ctr = self.buf(f"fps{i}.ctr", B, 3, M)
            call("p2pb_furthest_point_sampling", _p(coords[i]), B, Ns[i], M, _p(idx), _p(ctr), _vp(0), _s())
            coords.append(ctr)
            t = self.buf(f"bq{i}", B, M, L["K"], dtype=torch.int32)
            call("p2pb_ball_query", _p(coords[i + 1]), _p(coords[i]), B, M, Ns[i], _f(L["radius"]), L["K"], _p(t), _s())
            nidx.append(t)
            for P in L["pv"]:                       # voxel CSR of this level's PVConvs (level 0 is already cached)
                self.voxel_prep(preps, i, coords[i], P["r"])
            if i + 1 < n_levels:
                for P in W["sa"][i + 1]["pv"]:
                    self.voxel_prep(preps, i + 1, coords[i + 1], P["r"])
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            ev_level.append(ev)
        nn3 = []
        for j, L in enumerate(W["fp"]):
            lvl = L["lvl"]
            ix = self.buf(f"nn{j}.idx", B, 3, Ns[lvl], dtype=torch.int32)
            ww = self.buf(f"nn{j}.w", B, 3, Ns[lvl])
            call("p2pb_three_nn", _p(coords[lvl]), _p(coords[lvl + 1]), B, Ns[lvl], Ns[lvl + 1], _p(ix), _p(ww), _s())
            nn3.append((ix, ww))
        for i, L in enumerate(W["sa"]):
            for P in L["pv"]:
                self.voxel_prep(preps, i, coords[i], P["r"])
        for L in W["fp"]:
            for P in L["pv"]:
                self.voxel_prep(preps, L["lvl"], coords[L["lvl"]], P["r"])
        return coords, nidx, nn3, preps, ev_prep0, ev_level

    # ------------------------------------------------------------------------------------------------ one evaluation
    def prepare_cond(self, x_cond):
        """Step-invariant part: embed_feats(x_cond) (unet_pvc.py:184-188) when conditioning features are given."""
        if self.extra == 0:
            return
        B, N, W, fe = self.B, self.N, self.W, self.fe
        kc = pad32(self.extra)
        xc = self.buf("xcond.rows", B * N, kc)
        xc[:, :self.extra] = x_cond.permute(0, 2, 1).reshape(B * N, self.extra)   # layout change of the INPUT, once per call
        F0 = self.buf("F0", B * N, pad32(fe + 3))
        if "ef0" in W:
            raw, st, tl = self.gemm("ef0", [xc], [kc], W["ef0"], W["ef0b"], fe, N)
            A, Bc, _ = self.coef("ef0", st, tl, W["efn"], fe, N)
            h = self.buf("ef0.act", B * N, pad32(fe))
            self.act(raw, A, Bc, N, fe, h, act=1)
            dense.gemm_rows([h], W["ef3"], W["ef3b"], out=F0[:, :fe], ks=[pad32(fe)])
        else:
            F0[:, :fe] = xc[:, :fe]

    def evaluate(self, xt, sin, bridge=None):
        """One network evaluation: xt [B,3,N], sin = sinusoid of the noise level, [B,E] (row stride may be 0: every sample of a
        sampling step shares the noise level) -> eps rows [B*N, 16] (cols 0..2).  bridge = (coef device pointer table row, clip,
        pred_x0 buffer or None): the bridge update xt <- p_posterior(...) runs inside the head kernel (xt is updated in place)."""
        B, N, W, fe, E = self.B, self.N, self.W, self.fe, self.E
        Ns = self.Ns
        n_levels = len(W["sa"])
        # ---- everything that depends only on the time embedding: embedf MLP + every temb fold, one launch
        temb = self.buf("temb", B, E)
        call("p2pb_step_vectors", _p(sin), int(sin.stride(0)), _p(W["tw0"]), _p(W["tb0"]), _p(W["tw2"]), _p(W["tb2"]), B, E,
             _p(W.get("fold_w")), self._fold_R, _p(W.get("fold_ptr")), _p(W.get("fold_stride")), _p(temb), _s())
        # ---- geometry: FPS chain, ball queries, 3-NN, voxel CSRs (coordinates only).  It is latency-bound (one CTA per
        # patch for FPS) and independent of the features, so it runs on a side stream concurrently with the feature
        # embedding / global PointNet / AdaGN GEMMs below (fork/join with events; captured into the same CUDA graph).
        main = torch.cuda.current_stream()
        side = self._side_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            coords, nidx, nn3, preps, ev_prep0, ev_level = self._geometry(xt)
            join = torch.cuda.Event()
            join.record(side)
        # ---- point rows of the raw coordinates
        X0 = self.buf("X0", B * N, 32)
        call("p2pb_coords_to_rows", _p(xt), _p(X0), B, N, 32, 0, _s())
        F0 = self.buf("F0", B * N, pad32(fe + 3))
        call("p2pb_coords_to_rows", _p(xt), _p(F0), B, N, int(F0.stride(0)), fe, _s())
        if self.extra == 0:
            if "ef0" in W:
                raw, st, tl = self.gemm("ef0", [X0], [32], W["ef0"], W["ef0b"], fe, N)
                A, Bc, _ = self.coef("ef0", st, tl, W["efn"], fe, N)
                h = self.buf("ef0.act", B * N, pad32(fe))
                self.act(raw, A, Bc, N, fe, h, act=1)
                dense.gemm_rows([h], W["ef3"], W["ef3b"], out=F0[:, :fe], ks=[pad32(fe)])
            else:
                raise NotImplementedError("extra_feature_channels == 0 without embed_feats")
        # ---- global PointNet -> cond [B, cond_dim] -> all AdaGN (factor, bias) vectors in one GEMM
        if "pnet" in W:
            G = W["pnet"]
            x, kx = X0, 32
            pdt = torch.float32
            if G[0]["w"].dtype == torch.float16:       # half operands: coords rows and every activation that feeds a GEMM
                pdt = torch.float16
                x, kx = self.buf("X0h", B * N, 64, dtype=pdt), 64
                call("p2pb_coords_to_rows_f16", _p(xt), _p(x), B, N, 64, 0, _s())
            padk = pad64 if pdt == torch.float16 else pad32
            g_half = None
            for j, L in enumerate(G):
                nm = f"pnet{j}"
                bias2 = None
                if j == 2:
                    bias2 = self.buf(nm + ".gb", B, L["cout"])
                    self.linear(g_half, L["w_g"], None, 0, bias2)
                # the global max-pools (pvcnn.py:923-926, 930-931) come from the GEMM epilogue's column (max, min): Swish is
                # unimodal, so max over points of Swish(GN(x)) is attained at the column's largest or smallest x.  The last
                # layer's [B*N, 1024] output is never written.
                mm = (j == 1 or j == 3) and N % 128 == 0
                raw, st, tl = self.gemm(nm, [x], [kx], L["w"], L["b"], L["cout"], N, bias2=bias2, minmax=mm,
                                        store=not (mm and j == 3))
                colmm = self.last_colmm
                A, Bc, _ = self.coef(nm, st, tl, L["n"], L["cout"], N)
                if mm:
                    g = self.buf(nm + ".gmax" if j == 1 else "cond", B, L["cout"])
                    call("p2pb_gmax_minmax", _p(colmm), tl, B, L["cout"], _p(A), _p(Bc), 1, _p(g), _s())
                    if j == 1:
                        g_half = g
                        out = self.buf(nm + ".act", B * N, padk(L["cout"]), dtype=pdt)
                        self.act(raw, A, Bc, N, L["cout"], out, act=1)
                    else:
                        out, cond = None, g
                elif j == 1:      # activation rows AND global max over points (pvcnn.py:923-926)
                    out = self.buf(nm + ".act", B * N, pad32(L["cout"]))
                    g_half = self.buf(nm + ".gmax", B, L["cout"])
                    self.act(raw, A, Bc, N, L["cout"], out, act=1, gmax=g_half)
                elif j == 3:    # only the global max is needed (pvcnn.py:930-931)
                    out = None
                    cond = self.buf("cond", B, L["cout"])
                    self.act(raw, A, Bc, N, L["cout"], None, act=1, gmax=cond)
                else:
                    out = self.buf(nm + ".act", B * N, padk(L["cout"]), dtype=pdt)
                    self.act(raw, A, Bc, N, L["cout"], out, act=1)
                x, kx = out, padk(L["cout"])
            self.emd_all = self.buf("emd_all", B, self._emd_ld)
            dense.gemm_rows([cond], W["emd_w"], W["emd_b"], out=self.emd_all)
        else:
            self.emd_all = None
        main.wait_event(ev_prep0)   # level-0 voxel CSRs: all the first PVConv needs from the geometry stream
        # ---- set abstraction
        feats = F0
        skips = []
        for i, L in enumerate(W["sa"]):
            skips.append(feats)
            n_pts = Ns[i]
            for k, P in enumerate(L["pv"]):
                prep = self.voxel_prep(preps, i, coords[i], P["r"])
                feats = self.pvconv(f"sa{i}.pv{k}", P, feats, coords[i], prep, temb, n_pts)
            main.wait_event(ev_level[i])    # centres + neighbour lists of this level (and the next level's voxel CSRs)
            M, K, cg = Ns[i + 1], L["K"], L["c_grp"]
            if L["proj"]:
                feats = self.sa_projected(f"sa{i}", L, feats, coords[i], coords[i + 1], nidx[i], temb, n_pts, M, K, cg)
                continue
            g16 = L["mlp"][0]["w"].dtype == torch.float16
            kg = pad64(cg + 3) if g16 else pad32(cg + 3)
            grp = self.buf(f"sa{i}.grp", B * M * K, kg, dtype=torch.float16 if g16 else torch.float32)
            call("p2pb_group_rows_f16" if g16 else "p2pb_group_rows", _p(feats), int(feats.stride(0)), cg, _p(coords[i]),
                 _p(coords[i + 1]), _p(nidx[i]), _p(grp), int(grp.stride(0)), B, n_pts, M, K, _s())
            feats = self.mlp_chain(f"sa{i}.mlp", L["mlp"], [grp], [kg], M * K, temb, final_pool=K)
        main.wait_event(join)       # 3-NN tables and the remaining voxel CSRs (feature propagation)
        # ---- bottleneck attention, no residual: LinearAttention (modules.py:165-194) or softmax Attention (modules.py:197-264)
        nb = Ns[-1]
        if "att_qkv" in W:
            c = feats.shape[1]
            qkv = self.buf("att.qkv", B * nb, W["att_qkv"].shape[0])
            dense.gemm_rows([feats], W["att_qkv"], None, out=qkv, ks=[pad32(c)])
            hid = W["att_out"].shape[1]
            att = self.buf("att.ctx", B * nb, hid)
            call("p2pb_attention_small" if W["att_kind"] == "linear" else "p2pb_attention_softmax_small", _p(qkv), int(qkv.stride(0)),
                 B, W["heads"], nb, _p(att), hid, _s())
            fo = self.buf("att.out", B * nb, c)
            dense.gemm_rows([att], W["att_out"], W["att_outb"], out=fo, ks=[hid])
            feats = fo
        # ---- feature propagation
        for j, L in enumerate(W["fp"]):
            lvl = L["lvl"]
            n_up, n_low, c_low = Ns[lvl], Ns[lvl + 1], L["c_low"]
            ix, ww = nn3[j]
            itp = self.buf(f"fp{j}.itp", B * n_up, c_low)
            call("p2pb_interp_rows", _p(feats), int(feats.stride(0)), _p(ix), _p(ww), _p(itp), c_low, B, c_low, n_up, n_low, _s())
            skip = skips[lvl]
            feats = self.mlp_chain(f"fp{j}.mlp", L["mlp"], [itp, skip], [c_low, L["kp_skip"]], n_up, temb)
            for k, P in enumerate(L["pv"]):
                prep = self.voxel_prep(preps, lvl, coords[lvl], P["r"])
                feats = self.pvconv(f"fp{j}.pv{k}", P, feats, coords[lvl], prep, temb, n_up)
        # ---- classifier head (unet_pvc.py:147-154,263-267)
        cls = W["cls"]
        segs, ks = [feats], [pad32(feats.shape[1])]
        if len(cls) > 1:
            hmid = self.mlp_chain("cls", cls[:-1], segs, ks, N, temb)
            segs, ks = [hmid], [hmid.shape[1]]
        L = cls[-1]
        nm = f"cls.{len(cls) - 1}"
        raw, st, tl = self.gemm(nm, segs, ks, L["w"], L["b"], L["cout"], N)
        A, Bc, _ = self.coef(nm, st, tl, L["n"], L["cout"], N)
        eps = self.buf("eps", B * N, 16)
        # last GroupNorm + Swish, the C -> 3 projection (fp32) and, inside the sampling loop, the bridge update: one pass
        if bridge is None:
            call("p2pb_head_bridge", _p(raw), int(raw.stride(0)), _p(A), _p(Bc), _p(W["cls_out"]), _p(W["cls_outb"]), B, L["cout"], N,
                 _vp(0), _vp(0), 0, _vp(0), _vp(0), _p(eps), 16, _s())
        else:
            coef, clip, x0 = bridge
            call("p2pb_head_bridge", _p(raw), int(raw.stride(0)), _p(A), _p(Bc), _p(W["cls_out"]), _p(W["cls_outb"]), B, L["cout"], N,
                 _p(xt), _p(coef), int(bool(clip)), _p(xt), _p(x0), _p(eps), 16, _s())
        return eps

    def time_embedding(self, noise_level: float, out):
        """get_timestep_embedding + embedf (unet_pvc.py:156-169, 52-56) for a batch that shares one noise level."""
        import numpy as np

        half = self.E // 2
        e = np.log(10000) / (half - 1)
        freqs = torch.from_numpy(np.exp(np.arange(0, half) * -e)).float()
        ang = torch.tensor(noise_level, dtype=torch.float32) * freqs
        return torch.cat([torch.sin(ang), torch.cos(ang)]).to(self.dev)

    # ------------------------------------------------------------------------------------------------ sampling loop
    def _run_loop(self, pairs, log_set, clip, xs_buf, x0_buf):
        """All T steps on the current stream (this is what gets captured)."""
        B, N, E = self.B, self.N, self.E
        xt = self.buf("xt", B, 3, N)
        li = 0
        for s, (prev, step) in enumerate(pairs):
            sin = self.buf(f"temb.sin{len(pairs)}", len(pairs), E)
            row = sin[s:s + 1].expand(B, E)       # stride-0 view: every sample shares the noise level in sampling
            logged = prev in log_set
            x0 = x0_buf[li] if logged else None
            self.evaluate(xt, row, bridge=(self.buf(f"coef{len(pairs)}", len(pairs), 3)[s], clip, x0))
            if logged:
                xs_buf[li].copy_(xt)
                li += 1

    def prepare(self, x1, x_cond, pairs, log_steps):
        """Upload the per-step host tables and this call's inputs into the static buffers -> (log_set, xs_buf, x0_buf)."""
        B, N, E = self.B, self.N, self.E
        p = self.p2pb
        T = len(pairs)
        log_set = set(log_steps)
        n_log = sum(1 for prev, _ in pairs if prev in log_set)
        # per-step host tables: sinusoid of the noise level, posterior scalars (fp32, same op order as p_posterior); computed and
        # uploaded once per step grid (the tables of one T never change)
        sin = self.buf(f"temb.sin{T}", T, E)       # keyed by T: sample(steps=...) may change between calls
        coef = self.buf(f"coef{T}", T, 3)
        tkey = tuple(pairs)
        if getattr(self, "_tables_key", {}).get(T) != tkey:
            sin_h = torch.stack([self.time_embedding(float(p.noise_levels[step].item()), None) for _, step in pairs])
            coef_h = torch.tensor([p.posterior_coefs(prev, step) for prev, step in pairs], dtype=torch.float32)
            sin.copy_(sin_h)
            coef.copy_(coef_h.to(self.dev))
            if not hasattr(self, "_tables_key"):
                self._tables_key = {}
            self._tables_key[T] = tkey
        self.buf("xt", B, 3, N).copy_(x1.detach().to(self.dev, torch.float32))
        if x_cond is not None:
            self.prepare_cond(x_cond.detach().to(self.dev, torch.float32))
        xs_buf = self.buf(f"xs{n_log}", n_log, B, 3, N)
        x0_buf = self.buf(f"x0s{n_log}", n_log, B, 3, N)
        return log_set, xs_buf, x0_buf

    def sample(self, x1, x_cond, pairs, log_steps, clip):
        """pairs = [(prev_step, step)] in sampling order -> (xs, pred_x0s) [B, log_count, 3, N], logged steps flipped
        to ascending time like ``sample_ddpm`` (p2pb.py:215-262).  Runs on the engine's own device whatever the caller's
        current device is (denoise_object.py --gpu cuda:1)."""
        with torch.cuda.device(self.dev):
            return self._sample(x1, x_cond, pairs, log_steps, clip)

    def _sample(self, x1, x_cond, pairs, log_steps, clip):
        log_set, xs_buf, x0_buf = self.prepare(x1, x_cond, pairs, log_steps)
        xt = self.buf("xt", self.B, 3, self.N)
        key = (tuple(pairs), tuple(sorted(log_set)), bool(clip))
        if OPTIONS.no_graph:     # profiling aid: same kernels, no graph capture
            self._run_loop(pairs, log_set, clip, xs_buf, x0_buf)
            xs = torch.flip(xs_buf.permute(1, 0, 2, 3), dims=(1,)).clone()
            return xs, torch.flip(x0_buf.permute(1, 0, 2, 3), dims=(1,)).clone()
        entry = self._graphs.get(key)
        if entry is None:
            # first call: run eagerly once (allocates every buffer, sets kernel attributes), then capture
            l0 = launch_count()
            self._run_loop(pairs, log_set, clip, xs_buf, x0_buf)
            self.kernels_per_sample = launch_count() - l0
            torch.cuda.current_stream().synchronize()
            xt.copy_(x1.detach().to(self.dev, torch.float32))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_loop(pairs, log_set, clip, xs_buf, x0_buf)
            self._graphs[key] = (g,)
            xt.copy_(x1.detach().to(self.dev, torch.float32))
            entry = self._graphs[key]
        entry[0].replay()
        xs = torch.flip(xs_buf.permute(1, 0, 2, 3), dims=(1,)).clone()
        x0s = torch.flip(x0_buf.permute(1, 0, 2, 3), dims=(1,)).clone()
        return xs, x0s

    def half_overflows(self) -> int:
        """Values that did not fit IEEE half in any half-producing kernel since the last query (0 unless an activation
        exceeded 65504); waits for the current stream."""
        if not (self.halo_f16 or self.gemm_f16):
            return 0
        n = ctypes.c_uint(0)
        call("p2pb_half_overflow_count", 1, ctypes.byref(n), _s())
        return int(n.value)


class DualEngine:
    """n (default 2) part-batch engines whose T-step loops run as INDEPENDENT chains on separate streams inside one CUDA graph.

    Patches never interact (GroupNorm / SE / attention are per sample), so the chains need no synchronisation until the
    end of the call.  The evaluation is a strict chain of ~210 kernels of which ~90 are tiny, latency-bound launches
    (GroupNorm coefficients, per-sample linears, FPS, ...) that leave the GPU nearly idle; the big tensor-core kernels
    are persistent one-CTA-per-SM kernels that cannot overlap each other.  With several chains the small kernels of one
    part run in the shadow of another part's convolutions / GEMMs."""

    def __init__(self, p2pb, net, B: int, N: int, F: int, n_chains: int = 2, allow_half: bool = True):
        assert B % n_chains == 0
        self.B, self.N, self.F, self.n = B, N, F, n_chains
        self.part = B // n_chains
        self.halves = [Engine(p2pb, net, self.part, N, F, allow_half) for _ in range(n_chains)]
        self.tf32_layers = self.halves[0].tf32_layers
        self.dev = self.halves[0].dev
        self.dtype_name = self.halves[0].dtype_name
        self._graphs: Dict[tuple, tuple] = {}
        self._streams = [torch.cuda.Stream(device=self.dev) for _ in range(n_chains - 1)]
        self.kernels_per_sample = 0

    def half_overflows(self) -> int:
        return self.halves[0].half_overflows()

    def _run_all(self, pairs, prep, clip):
        """Fork: chain 0 on the current stream, the others on their own streams; join."""
        cur = torch.cuda.current_stream()
        for st in self._streams:
            st.wait_stream(cur)
        self.halves[0]._run_loop(pairs, prep[0][0], clip, prep[0][1], prep[0][2])
        for st, e, pr in zip(self._streams, self.halves[1:], prep[1:]):
            with torch.cuda.stream(st):
                e._run_loop(pairs, pr[0], clip, pr[1], pr[2])
        for st in self._streams:
            cur.wait_stream(st)

    def sample(self, x1, x_cond, pairs, log_steps, clip):
        with torch.cuda.device(self.dev):
            return self._sample(x1, x_cond, pairs, log_steps, clip)

    def _sample(self, x1, x_cond, pairs, log_steps, clip):
        h = self.part
        parts = [(x1[i * h:(i + 1) * h], None if x_cond is None else x_cond[i * h:(i + 1) * h]) for i in range(self.n)]
        prep = [e.prepare(x, c, pairs, log_steps) for e, (x, c) in zip(self.halves, parts)]
        key = (tuple(pairs), tuple(sorted(prep[0][0])), bool(clip))
        if OPTIONS.no_graph:
            self._run_all(pairs, prep, clip)
        else:
            if key not in self._graphs:
                l0 = launch_count()
                self._run_all(pairs, prep, clip)         # eager first run: allocates every buffer
                self.kernels_per_sample = launch_count() - l0
                torch.cuda.synchronize(self.dev)
                for e, (x, c) in zip(self.halves, parts):
                    e.buf("xt", h, 3, self.N).copy_(x.detach().to(self.dev, torch.float32))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run_all(pairs, prep, clip)
                self._graphs[key] = (g,)
                for e, (x, c) in zip(self.halves, parts):
                    e.buf("xt", h, 3, self.N).copy_(x.detach().to(self.dev, torch.float32))
            self._graphs[key][0].replay()
        xs = torch.cat([torch.flip(p[1].permute(1, 0, 2, 3), dims=(1,)) for p in prep], 0)
        x0s = torch.cat([torch.flip(p[2].permute(1, 0, 2, 3), dims=(1,)) for p in prep], 0)
        return xs, x0s


def get_engine(p2pb, net, x_shape, cond_shape, allow_dual: bool = False, allow_half: bool = True):
    """Engine for (net, B, N, F), built once.  allow_dual (the sampling loop): with P2PB_CHAINS=n the batch is split into n
    independent part-batch chains (DualEngine).  Default 1: measured on B200 (PVDS) one chain 366 patches/s, two chains 361
    at 64 patches, 390 vs 384 at 128 -- since the small kernels were rewritten, batch efficiency of the big kernels
    outweighs the overlap."""
    B, _, N = x_shape
    F = 0 if cond_shape is None else cond_shape[1]
    n = int(OPTIONS.chains)
    while n > 1 and (B % n != 0 or B // n < 8):
        n -= 1
    dual = allow_dual and n > 1
    # the packed weights are a snapshot: key them by the parameters' in-place version counters, so that a
    # load_state_dict / optimizer step after the first sample() rebuilds the engine instead of running stale weights
    version = sum(int(p._version) for p in net.parameters()) + sum(int(b._version) for b in net.buffers())
    allow_half = allow_half and id(net) not in getattr(p2pb, "_no_half", ())
    key = (id(net), B, N, F, n if dual else 1, allow_half)
    entry = p2pb._engines.get(key)
    if entry is not None and entry[0] != version:
        for k in [k for k in p2pb._engines if k[0] == id(net)]:
            del p2pb._engines[k]            # every shape's engine of this net is stale: free their buffers
        entry = None
    if entry is None:
        dev = next(net.parameters()).device
        with torch.cuda.device(dev):
            eng = DualEngine(p2pb, net, B, N, F, n, allow_half) if dual else Engine(p2pb, net, B, N, F, allow_half)
        entry = (version, eng)
        p2pb._engines[key] = entry
    eng = entry[1]
    p2pb.last_engine = eng
    return eng
