"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol include/p2pb_b200.h declares;
the drop-in modules expose the reference's names; the product path refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "p2pb_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(p2pb_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from p2pb_b200 import build

    lib = ctypes.CDLL(build.build())
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/p2pb_b200.h but not exported"
    lib.p2pb_last_error.restype = ctypes.c_char_p
    assert lib.p2pb_abi_version() >= 1
    assert lib.p2pb_last_error() == b""


def test_dropin_module_has_reference_names():
    from p2pb_b200 import _pvcnn_backend, pointnet2_batch_cuda as m

    # pointnet2_api.cpp:31-47
    for n in ["avg_voxelize_forward", "avg_voxelize_backward", "trilinear_devoxelize_forward",
              "trilinear_devoxelize_backward", "ball_query", "three_nearest_neighbors_interpolate_forward",
              "three_nearest_neighbors_interpolate_backward", "grouping_forward", "grouping_backward",
              "gather_features_forward", "gather_features_backward", "furthest_point_sampling_forward"]:
        assert callable(getattr(m, n)), n
    assert callable(_pvcnn_backend.furthest_point_sampling)  # third_party/pvcnn/functional/src/bindings.cpp:15
    with pytest.raises(NotImplementedError):
        m.grouping_backward(None, None, 0)


def test_no_cpu_fallback():
    from p2pb_b200 import ops
    from p2pb_b200._lib import P2PBError

    with pytest.raises(P2PBError):
        ops.furthest_point_sampling(torch.zeros(1, 3, 16), 4)
    with pytest.raises(P2PBError):
        ops.ball_query(torch.zeros(1, 3, 4), torch.zeros(1, 3, 16), 0.1, 8)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under p2pb_b200/ (nor the entry scripts) may reference it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "p2pb_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|from\s+\.\.?oracle|p2pb_oracle\.so", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_seeded_weights_equal_oracle_weights():
    """bench/smoke weights come from the product's own generator; it must equal the oracle's (same per-key streams)."""
    from oracle import model as OM
    from p2pb_b200.config import load_yaml
    from p2pb_b200.model_loader import seeded_state_dict
    from p2pb_b200.unet_pvc import PVCNN2Unet

    cfg = load_yaml(os.path.join(ROOT, "p2pb_b200", "configs", "PVDS_PUNet.yaml"))
    a = seeded_state_dict(PVCNN2Unet(cfg), seed=0)
    b = OM.make_state_dict(cfg.to_dict(), seed=0)
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_header_declares_every_exported_symbol():
    """include/p2pb_b200.h is the complete boundary: every P2PB_API function in csrc/ is declared there."""
    import glob

    exported = set()
    for f in glob.glob(os.path.join(ROOT, "p2pb_b200", "csrc", "*.cu")):
        exported |= set(re.findall(r"P2PB_API\s+[\w\s\*]+?\b(p2pb_\w+)\s*\(", open(f).read()))
    assert exported == set(declared_symbols()), exported ^ set(declared_symbols())
