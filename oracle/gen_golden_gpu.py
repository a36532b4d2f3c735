"""Generate ``tests/golden/ref_ops.npz`` by running the REFERENCE's own compiled CUDA kernels on a B200.

TEST INFRASTRUCTURE ONLY.  Needs ``oracle/_ref/*.so`` (built in the container by ``oracle/build_ref.py`` from the
sources under /root/reference; the .so files travel to the GPU box with the gpurun snapshot).  Run once per change of
the cases:   gpurun -- python -m oracle.gen_golden_gpu   -> gpurun_out/ref_ops.npz  -> copy to tests/golden/.

Every op of the hot path gets seeded inputs (stored) and the reference kernel's outputs (stored).  Includes
tie-heavy inputs (coordinates snapped to a coarse lattice) so FPS / ball-query / 3-NN tie-breaks are pinned, and
the avg_voxelize output of two runs (the reference's own run-to-run spread from fp32 atomics).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def load_ext(name):
    path = os.path.join(HERE, "_ref", name + ".so")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cloud(B, N, seed, snap=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, N, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * (1 + 0.05 * torch.randn(B, 1, N, generator=g))
    x = x * 0.5
    if snap:
        x = torch.round(x * snap) / snap
    return x.contiguous()


def main():
    ref = load_ext("pointnet2_batch_cuda")
    cham = load_ext("chamfer_3D")
    dev = torch.device("cuda:0")
    out = {}
    for tag, snap in (("a", None), ("tie", 8.0)):
        B, N, M, C, r = 2, 2048, 512, 6, 16
        coords = cloud(B, N, 10 if snap is None else 11, snap)
        g = torch.Generator().manual_seed(5)
        feats = torch.randn(B, C, N, generator=g)
        cd, fd = coords.to(dev), feats.to(dev)
        idx = ref.furthest_point_sampling_forward(cd, M)
        centers = ref.gather_features_forward(cd, idx)
        nidx = ref.ball_query(centers, cd, 0.1, 32)
        grouped = ref.grouping_forward(fd, nidx)
        cfeat = torch.randn(B, C, M, generator=g).to(dev)
        interp, iidx, iw = ref.three_nearest_neighbors_interpolate_forward(cd, centers, cfeat)
        # voxel coords the way models/pvcnn.py:215-231 makes them (torch ops on the GPU)
        nc = cd - cd.mean(2, keepdim=True)
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0) + 0.5
        nc = torch.clamp(nc * r, 0, r - 1)
        vox = torch.round(nc).to(torch.int32)
        vo, vind, vcnt = ref.avg_voxelize_forward(fd, vox, r)
        vo2, _, _ = ref.avg_voxelize_forward(fd, vox, r)
        grid = torch.randn(B, C, r ** 3, generator=g).to(dev)
        dv, _, _ = ref.trilinear_devoxelize_forward(r, False, nc.contiguous(), grid)
        p1 = cd.transpose(1, 2).contiguous()
        p2 = (cd[:, :, : N // 2] + 0.01).transpose(1, 2).contiguous()
        d1 = torch.zeros(B, N, device=dev); d2 = torch.zeros(B, N // 2, device=dev)
        i1 = torch.zeros(B, N, dtype=torch.int32, device=dev); i2 = torch.zeros(B, N // 2, dtype=torch.int32, device=dev)
        cham.forward(p1, p2, d1, d2, i1, i2)
        torch.cuda.synchronize()
        c = lambda t: t.detach().cpu().numpy()
        out.update({
            f"{tag}_coords": c(coords), f"{tag}_feats": c(feats), f"{tag}_cfeat": c(cfeat), f"{tag}_grid": c(grid),
            f"{tag}_fps_idx": c(idx), f"{tag}_centers": c(centers), f"{tag}_ball_idx": c(nidx),
            f"{tag}_grouped": c(grouped).astype(np.float32), f"{tag}_interp": c(interp), f"{tag}_interp_idx": c(iidx),
            f"{tag}_interp_w": c(iw), f"{tag}_norm_coords": c(nc), f"{tag}_vox": c(vox), f"{tag}_vox_out": c(vo),
            f"{tag}_vox_out2": c(vo2), f"{tag}_vox_ind": c(vind), f"{tag}_vox_cnt": c(vcnt), f"{tag}_devox": c(dv),
            f"{tag}_cd_d1": c(d1), f"{tag}_cd_d2": c(d2), f"{tag}_cd_i1": c(i1), f"{tag}_cd_i2": c(i2),
        })
    # small-N FPS chain like the deep U-Net levels (N < 512 exercises the "thread without a point" slots)
    coords = cloud(3, 128, 12)
    idx = ref.furthest_point_sampling_forward(coords.to(dev), 32)
    out["small_coords"], out["small_fps_idx"] = coords.numpy(), idx.cpu().numpy()
    # EMD known answer of the reference (metrics/PyTorchEMD/test_emd_loss.py:6-20) + a random case
    emd = load_ext("emd_cuda")
    p1 = torch.tensor([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]]).repeat(3, 1, 1).to(dev)
    p2 = torch.tensor([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]]).repeat(3, 1, 1).to(dev)
    match = emd.approxmatch_forward(p1, p2)
    cost = emd.matchcost_forward(p1, p2, match)
    out["emd_known_cost"] = cost.cpu().numpy()
    a = cloud(2, 256, 13).transpose(1, 2).contiguous().to(dev)
    b = (cloud(2, 256, 14) * 1.1).transpose(1, 2).contiguous().to(dev)
    match = emd.approxmatch_forward(a, b)
    out["emd_a"], out["emd_b"] = a.cpu().numpy(), b.cpu().numpy()
    out["emd_cost"] = emd.matchcost_forward(a, b, match).cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ref_ops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
