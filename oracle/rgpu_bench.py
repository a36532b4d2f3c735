"""R-GPU arm: throughput of the UNMODIFIED reference (its ``models/*.py`` from ``baseline/_ref`` over its own compiled
``oracle/_ref/pointnet2_batch_cuda.so``) on this box's GPU, for the bench's workloads.  BASELINE INFRASTRUCTURE ONLY: a
separate process started by ``bench.py`` (never imported by the product), PyTorch defaults as the reference runs them (cuDNN
convolutions in TF32, matmul fp32), seeded weights and synthetic patches identical to the B200 arm.

Prints ONE JSON line:  {"available": true, "pvds_2048": {...}, "pvdl_8192_xyz": {...}}   or   {"available": false, "why": "..."}

Usage:  python -m oracle.rgpu_bench [--device cuda:0] [--pvdl]
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--pvdl", action="store_true", help="also time PVDL N=8192 xyz (BASELINE config 3), 32 patches")
    args = ap.parse_args()
    try:
        import torch

        from oracle import model as OM
        from oracle.ref_import import AttrDict, import_reference, load_ref_cfg, load_ref_extension, reference_root

        ref = reference_root()
        if ref is None:
            raise RuntimeError("no reference tree (baseline/_ref snapshot missing: run oracle/snapshot_ref.py in the build container)")
        if not os.path.exists(os.path.join(HERE, "_ref", "pointnet2_batch_cuda.so")):
            raise RuntimeError("oracle/_ref/pointnet2_batch_cuda.so missing (oracle/build_ref.py)")
        import bench
        from tests.helpers import patch_input

        ext = load_ref_extension("pointnet2_batch_cuda")
        PVCNN2Unet, P2PB = import_reference(ext, ref)
        torch.set_grad_enabled(False)
        dev = args.device
        torch.cuda.set_device(dev)

        def run(cfg, x, T):
            acfg = AttrDict.wrap(copy.deepcopy(cfg))
            acfg.gpu = dev
            acfg.model.ema = False
            net = PVCNN2Unet(acfg)
            net.load_state_dict(OM.make_state_dict(cfg, seed=0), strict=True)
            model = P2PB(acfg, net)
            model.eval()
            x = x.to(dev)
            model.sample(x_start=x, steps=2, log_count=1, verbose=False, use_ema=False)       # warm-up (cuDNN autotune)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.sample(x_start=x, steps=T, log_count=1, verbose=False, use_ema=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            return {"patches": int(x.shape[0]), "npoints": int(x.shape[2]), "T": T, "ms_per_sample_call": ms,
                    "patches_per_s": x.shape[0] / (ms / 1e3)}

        out = {"available": True, "what": "unmodified reference models/*.py over its own pointnet2_batch_cuda (sm_100), "
                                         "PyTorch defaults (cuDNN TF32 convolutions)", "gpu": torch.cuda.get_device_name(0)}
        out["pvds_2048"] = run(load_ref_cfg("PVDS_PUNet", ref), bench.synth_patches(64, 2048, seed=1000), 30)
        if args.pvdl:
            cfgL = load_ref_cfg("PVDL_SNPP", ref, **{"data.npoints": 8192, "model.extra_feature_channels": 0})
            out["pvdl_8192_xyz"] = run(cfgL, patch_input(32, 8192, seed=7), 30)
        print(json.dumps(out), flush=True)
    except Exception as ex:  # the bench line must not depend on the baseline being runnable
        print(json.dumps({"available": False, "why": repr(ex)[:300]}), flush=True)


if __name__ == "__main__":
    main()
