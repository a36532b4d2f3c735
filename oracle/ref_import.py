"""Import the REFERENCE's real Python model code (``models/{unet_pvc,pvcnn,modules,p2pb}.py`` + its six op wrapper files)
from a reference tree, with the op extension of the caller's choice.  TEST / BASELINE INFRASTRUCTURE ONLY.

The recipe is SURVEY.md App. C: empty package modules instead of the heavy ``third_party/openpoints/__init__.py`` chain, three
stubs for absent pure-Python dependencies (``ema_pytorch``, ``omegaconf``, ``emd_assignment``).  The tree is either
``/root/reference`` (build container) or the git-ignored snapshot ``baseline/_ref`` made by ``oracle/snapshot_ref.py`` (the
only copy that exists on the GPU box).  ``ops_module`` is what the wrappers see as ``pointnet2_cuda``:
  * ``oracle.ops``                                  -> the reference's model code on CPU (R-CPU; ``oracle/gen_golden.py``)
  * ``oracle/_ref/pointnet2_batch_cuda.so``         -> the unmodified reference on a GPU (R-GPU; ``oracle/gen_golden_rgpu.py``)
  * ``p2pb_b200.pointnet2_batch_cuda``              -> the reference's model code over this repo's op shim (drop-in test)
"""
from __future__ import annotations

import copy
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_root() -> str | None:
    """/root/reference when present (build container), else the snapshot that travels to the GPU box."""
    for p in (os.environ.get("P2PB_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if p and os.path.isdir(os.path.join(p, "models")):
            return p
    return None


class AttrDict(dict):
    """attribute + ``in`` + ``.get`` access, what the reference needs from an OmegaConf node."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__

    @staticmethod
    def wrap(d):
        if isinstance(d, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in d.items()})
        if isinstance(d, list):
            return [AttrDict.wrap(v) for v in d]
        return d


def load_ref_extension(name: str = "pointnet2_batch_cuda"):
    """The reference's own compiled extension (oracle/build_ref.py -> oracle/_ref/<name>.so)."""
    path = os.path.join(HERE, "_ref", name + ".so")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def import_reference(ops_module, ref: str | None = None):
    """-> (PVCNN2Unet, P2PB) classes of the reference, wired to ``ops_module``."""
    import torch

    ref = ref or reference_root()
    assert ref is not None, "no reference tree (/root/reference or baseline/_ref)"

    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k.startswith("third_party")]:
        del sys.modules[k]          # a previous import with another ops module / tree
    pkg("third_party", f"{ref}/third_party")
    pkg("third_party.openpoints", f"{ref}/third_party/openpoints")
    pkg("third_party.openpoints.models", f"{ref}/third_party/openpoints/models")
    pkg("third_party.openpoints.cpp", f"{ref}/third_party/openpoints/cpp").pointnet2_cuda = ops_module
    layers = pkg("third_party.openpoints.models.layers", f"{ref}/third_party/openpoints/models/layers")
    for sub, name in [("voxelization", "avg_voxelize"), ("devoxelization", "trilinear_devoxelize"),
                      ("ball_query", "ball_query"), ("interpolatation", "nearest_neighbor_interpolate"),
                      ("sampling", "furthest_point_sample_pvcnn"), ("group", "pvcnn_grouping")]:
        full = f"third_party.openpoints.models.layers.{sub}"
        spec = importlib.util.spec_from_file_location(full, f"{ref}/third_party/openpoints/models/layers/{sub}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        setattr(layers, name, getattr(mod, name))
    ema = types.ModuleType("ema_pytorch")

    class EMA(torch.nn.Module):
        def __init__(self, model, beta=0.999):
            super().__init__()
            self.ema_model = copy.deepcopy(model)

        def forward(self, *a, **k):
            return self.ema_model(*a, **k)

    ema.EMA = EMA
    sys.modules["ema_pytorch"] = ema
    oc = types.ModuleType("omegaconf")
    oc.DictConfig = dict
    oc.OmegaConf = object
    sys.modules["omegaconf"] = oc
    sys.modules["emd_assignment"] = types.ModuleType("emd_assignment")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from models.p2pb import P2PB  # noqa
    from models.unet_pvc import PVCNN2Unet  # noqa

    return PVCNN2Unet, P2PB


def load_ref_cfg(name: str, ref: str | None = None, **over) -> dict:
    """The reference's own YAML + dotted-key overrides."""
    import yaml

    ref = ref or reference_root()
    cfg = yaml.safe_load(open(f"{ref}/configs/{name}.yaml"))
    for k, v in over.items():
        node = cfg
        ks = k.split(".")
        for kk in ks[:-1]:
            node = node[kk]
        node[ks[-1]] = v
    return cfg
