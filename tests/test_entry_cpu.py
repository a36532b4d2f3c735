"""CPU tests of the entry-point plumbing: CLI parity with the reference scripts, opt.yaml discovery, PLY/XYZ I/O,
checkpoint format."""
import os

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ckpt_dir(tmp_path):
    cfg = yaml.safe_load(open(os.path.join(ROOT, "p2pb_b200", "configs", "PVDS_PUNet.yaml")))
    (tmp_path / "opt.yaml").write_text(yaml.safe_dump(cfg))
    return str(tmp_path / "step_0.pth")


def test_denoise_object_cli_flags_and_config_discovery(tmp_path):
    import denoise_object as D

    cfg = D.parse_args(["--data_path", "a.xyz", "--save_path", "b.xyz", "--model_path", _ckpt_dir(tmp_path), "--steps", "7", "--k", "2"])
    assert cfg.steps == 7 and cfg.k == 2 and cfg.use_ema is False and cfg.gpu == "cuda:0" and cfg.seed == 42
    assert cfg.model.PVD.channels == [32, 64, 128, 256, 512] and cfg.diffusion.ot_ode is True
    assert cfg.restart is False and cfg.local_rank == 0


def test_denoise_room_cli_flags(tmp_path):
    import denoise_room as D

    cfg = D.parse_args(["--room_path", "r.ply", "--model_path", _ckpt_dir(tmp_path)])
    assert cfg.steps == 5 and cfg.k == 4 and cfg.batch_size == 32 and cfg.use_ema is True and cfg.average_predictions is True
    assert cfg.feature_name == "dino_iphone" and cfg.gpu == "cuda:0"


def test_ply_roundtrip_and_xyz_format(tmp_path):
    from p2pb_b200.io_ply import read_ply, write_array_to_xyz, write_ply

    g = np.random.default_rng(0)
    pts, col = g.normal(size=(100, 3)), g.integers(0, 256, size=(100, 3)) / 255.0
    write_ply(str(tmp_path / "a.ply"), pts, col)
    p2, c2 = read_ply(str(tmp_path / "a.ply"))
    np.testing.assert_array_equal(p2, pts)
    np.testing.assert_allclose(c2, col, atol=1e-9)
    write_array_to_xyz(str(tmp_path / "a.xyz"), pts[:3])
    txt = (tmp_path / "a.xyz").read_text()
    assert txt == "\n".join(" ".join("%8f" % v for v in row) for row in pts[:3])      # utils/utils.py:5-10 format
    np.testing.assert_allclose(np.loadtxt(str(tmp_path / "a.xyz")), pts[:3], atol=1e-6)


def test_checkpoint_format_roundtrip(tmp_path):
    """{"step","model_state","optimizer_state"} with model.* / ema.* keys (train.py:169-174, model_loader.py:116-159)."""
    from p2pb_b200.config import load_yaml
    from p2pb_b200.model_loader import extract_from_state_dict, save_checkpoint, seeded_state_dict
    from p2pb_b200.p2pb import P2PB
    from p2pb_b200.unet_pvc import PVCNN2Unet

    cfg = load_yaml(os.path.join(ROOT, "p2pb_b200", "configs", "PVDS_PUNet.yaml"))
    cfg.gpu = "cpu"
    net = PVCNN2Unet(cfg)
    net.load_state_dict(seeded_state_dict(net, 3))
    model = P2PB(cfg, net)
    path = str(tmp_path / "step_5.pth")
    save_checkpoint(path, model, step=5)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"step", "model_state", "optimizer_state"} and ck["step"] == 5
    ms = ck["model_state"]
    assert "model.sa_layers.0.0.voxel_layers.0.weight" in ms and any(k.startswith("ema.ema_model.") for k in ms)
    assert not any(k in ms for k in ("betas", "std_fwd", "loss_weight"))       # schedule tables are not persisted
    net2 = PVCNN2Unet(cfg)
    net2.load_state_dict(extract_from_state_dict(ms, "model."))
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    # a checkpoint saved under DP/DDP has model.module.* keys: both prefixes must resolve (model_loader.py:125-130)
    wrapped = {k.replace("model.", "model.module.", 1) if k.startswith("model.") else k: v for k, v in ms.items()}
    assert set(extract_from_state_dict(wrapped, "model.module.")) == set(net.state_dict())
