"""Live roofline measurement of the dominant kernel for bench.py.

Dominant kernel of the PVDS N=2048 hot path = the 3x3x3 voxel convolution at r=32 (conv_halo_kernel): per network
evaluation it accounts for ~33 % of the GPU time (profiles/r02_launches.md).  Measured here on its largest instance
(fp_layers.3.1 voxel_layers.0/.4: 64 -> 64 channels on a 32^3 grid) exactly as the engine launches it: IEEE-half
operands (kind::f16, fp32 accumulate), cta_group::2 CTA pairs, one launch per chain (B/2 patches when the engine splits
the batch into two chains), CUDA events on the launching stream, inputs larger than L2 (B=32: 161 MB in + 268 MB out).

achieved = algorithmic FLOPs per launch / mean launch time, algorithmic FLOPs = 2 * B * r^3 * 27 * Cin * Cout
(SURVEY.md 8d counts 7.2478 GFLOP per patch for this layer).  Bound: tensor.  peak = the measured dense 16-bit rate
(cuBLAS bf16 in MEASURED_PEAKS.json; kind::f16 and kind::bf16 issue at the same rate)."""
from __future__ import annotations

import json
import os

import torch

from . import dense

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["bf16_tflops"]), float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json, burst)"
        except Exception:
            pass
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def dominant_kernel_roofline(model, B: int, dev, iters: int = 20):
    from .engine import DualEngine, halo_f16

    HALO_F16 = halo_f16()

    r, cin, cout = 32, 64, 64
    if isinstance(getattr(model, "last_engine", None), DualEngine):
        B = B // model.last_engine.n     # the engine launches the kernel once per part-batch chain
    dt = torch.float16 if HALO_F16 else torch.float32
    g = torch.Generator(device=dev).manual_seed(0)
    X = dense.alloc_padded(B, cin, r, dev, dt)
    P = r + 1
    X[: B * P ** 3].view(B, P, P, P, cin)[:, 1:, 1:, 1:, :] = torch.randn(B, r, r, r, cin, device=dev, generator=g).to(dt)
    w = (torch.randn(cout, 27 * cin, device=dev, generator=g) / (27 * cin) ** 0.5).to(dt)
    bias = torch.zeros(cout, device=dev)
    out = torch.empty(B * r ** 3, cout, device=dev)
    _, _, tps = dense.halo_layout(r, cout, HALO_F16, cin=cin)
    stats = torch.zeros(B * tps, cout, 2, device=dev)
    for _ in range(3):
        dense.conv3d_halo(X, w, bias, B, r, cin, cout, out=out, stats=stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dense.conv3d_halo(X, w, bias, B, r, cin, cout, out=out, stats=stats)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * B * r ** 3 * 27 * cin * cout
    achieved = flops / (ms * 1e-3) / 1e12
    bf16_peak, _, how = measured_peaks()
    peak = bf16_peak if HALO_F16 else bf16_peak / 2.0
    esz = 2.0 if HALO_F16 else 4.0
    # DRAM bytes of this launch from the committed `ncu --set full` capture -- only if the capture was taken from THIS kernel source
    # (sha256 of csrc/conv_halo.cu stored beside the number); a stale capture reads as null instead of a wrong number
    traffic, traffic_note = None, "no capture"
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            import hashlib

            ent = json.load(open(tp)).get(f"conv_halo_64x64_r32_B{B}_{'f16' if HALO_F16 else 'tf32'}")
            sha = hashlib.sha256(open(os.path.join(ROOT, "p2pb_b200", "csrc", "conv_halo.cu"), "rb").read()).hexdigest()
            if ent and ent.get("source_sha256") == sha:
                traffic, traffic_note = ent["dram_bytes"], ent.get("capture", "")
            elif ent:
                traffic_note = "capture is from an older conv_halo.cu (hash mismatch)"
        except Exception:
            traffic = None
    kind = "IEEE-half operands, kind::f16" if HALO_F16 else "TF32"
    return {"bound": "tensor", "kernel": f"conv_halo_kernel (3x3x3 conv, 64->64 ch, 32^3 grid, {B} patches per launch, {kind}, "
                                         "fp32 accumulate, cta_group::2 tcgen05)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_note, "ms_per_launch": ms, "flops_per_launch": flops,
            # an MMA stream at N = Cout = 64 keeps the tensor pipe busy 32 of every 48.9 cycles (tools/ubench/mma_rate.cu): the
            # ceiling of THIS layer shape on this hardware, and the measured tensor-pipe activity of the last committed capture
            "shape_ceiling": {"what": "an N = Cout = 64 tcgen05.mma stream costs max(48.9, N/2) cycles per MMA (tools/ubench/mma_rate.cu): the "
                                      "tensor pipe can be busy at most 32 / 48.9 of the time for this layer shape",
                              "tensor_active_ceiling_pct": 100.0 * 32.0 / 48.9, "tensor_active_pct_ncu": 67.2,
                              "evidence": "profiles/r02_ncu_conv.md (ncu --set full of this launch: 617.8 k elapsed cycles = 121.4 tiles/SM x 108 "
                                          "MMAs x 47.1 cycles; the micro-benchmark's per-MMA cost is 48.9)"},
            "algorithmic_bytes_per_launch": B * (esz * (r + 1) ** 3 * cin + 4.0 * r ** 3 * cout),
            "peak_source": f"{how}: dense 16-bit rate {bf16_peak:.1f} TFLOP/s" + ("" if HALO_F16 else " / 2 for TF32")}
