"""Snapshot the few reference files the R-GPU baseline needs into the git-ignored ``baseline/_ref/`` so that they travel to
the GPU box (``/root/reference`` does not exist there).  TEST / BASELINE INFRASTRUCTURE ONLY: nothing under ``baseline/_ref`` is
tracked, nothing in the product imports it; it is the reference-arm install location of the bench contract.

Copied verbatim (no edits): ``models/*.py``, the six op wrappers ``third_party/openpoints/models/layers/{voxelization,
devoxelization,ball_query,interpolatation,sampling,group}.py``, ``metrics/__init__.py`` + ``metrics/emd_assignment/
{__init__,emd_module}.py`` (import chain of ``models/loss.py:8``), ``configs/*.yaml``, ``test.xyz``.

Usage:  python oracle/snapshot_ref.py        (in the build container; idempotent)
"""
from __future__ import annotations

import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("P2PB_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = (
    ["models/*.py", "configs/*.yaml", "test.xyz", "metrics/__init__.py", "metrics/emd_assignment/__init__.py",
     "metrics/emd_assignment/emd_module.py"]
    + [f"third_party/openpoints/models/layers/{n}.py"
       for n in ("voxelization", "devoxelization", "ball_query", "interpolatation", "sampling", "group")]
)


def main() -> int:
    if not os.path.isdir(os.path.join(REF, "models")):
        print(f"[snapshot_ref] {REF} not present; keeping {DST} as it is")
        return 0
    n = 0
    for pat in FILES:
        for src in sorted(glob.glob(os.path.join(REF, pat))):
            rel = os.path.relpath(src, REF)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            n += 1
    print(f"[snapshot_ref] {n} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
