"""Build the REFERENCE's own native extensions into ``oracle/_ref/`` (test infrastructure, never shipped).

This is the "strengthen the oracle with the real reference" recipe: the reference's CUDA
sources are compiled *where they lie* under ``/root/reference`` (nothing is copied into the repo)
with the same flags as the reference's ``setup.py`` (``-O2``, arch from TORCH_CUDA_ARCH_LIST;
``third_party/openpoints/cpp/pointnet2_batch/setup.py:32``), but by this short script instead of the
reference's build system.  Outputs go only to ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored, so
the ``.so`` files travel to the GPU box, where ``/root/reference`` does not exist).

Built modules (all CUDA-only, so they can only *run* on the GPU box):
  * ``pointnet2_batch_cuda``  - the 7 PVCNN forward ops the hot path calls
                                (``third_party/openpoints/cpp/pointnet2_batch/src/pointnet2_api.cpp:15-47``)
  * ``chamfer_3D``            - parity metric (``metrics/chamfer3D/chamfer_cuda.cpp``)
  * ``emd_cuda``              - parity metric (``metrics/PyTorchEMD/cuda/emd.cpp``)

Usage:  python oracle/build_ref.py [--only pointnet2_batch_cuda]
"""
from __future__ import annotations

import argparse
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("P2PB_REFERENCE_ROOT", "/root/reference")

MODULES = {
    "pointnet2_batch_cuda": {
        "dir": "third_party/openpoints/cpp/pointnet2_batch/src",
        "patterns": ["*.cpp", "*.cu"],
    },
    "chamfer_3D": {"dir": "metrics/chamfer3D", "patterns": ["chamfer_cuda.cpp", "chamfer3D.cu"]},
    "emd_cuda": {"dir": "metrics/PyTorchEMD/cuda", "patterns": ["emd.cpp", "emd_kernel.cu"]},
}


def build(name: str, verbose: bool = False) -> str | None:
    from torch.utils import cpp_extension

    spec = MODULES[name]
    src_dir = os.path.join(REF, spec["dir"])
    if not os.path.isdir(src_dir):
        print(f"[build_ref] {src_dir} not present - skipping {name}")
        return None
    sources = []
    for pat in spec["patterns"]:
        sources += sorted(glob.glob(os.path.join(src_dir, pat)))
    build_dir = os.path.join(OUT, "build_" + name)
    os.makedirs(build_dir, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    cpp_extension.load(
        name=name,
        sources=sources,
        extra_cflags=["-O2", "-w"],
        extra_cuda_cflags=["-O2", "-w"],
        extra_include_paths=[src_dir],
        build_directory=build_dir,
        verbose=verbose,
        is_python_module=True,
        with_cuda=True,
    )
    so = os.path.join(build_dir, name + ".so")
    dst = os.path.join(OUT, name + ".so")
    shutil.copyfile(so, dst)
    # keep only the .so (objects are large and need not travel)
    shutil.rmtree(build_dir, ignore_errors=True)
    print(f"[build_ref] built {dst}")
    return dst


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    names = [args.only] if args.only else list(MODULES)
    for n in names:
        dst = os.path.join(OUT, n + ".so")
        if os.path.exists(dst) and not args.force:
            print(f"[build_ref] {dst} exists")
            continue
        build(n, args.verbose)
    return 0


if __name__ == "__main__":
    sys.exit(main())
