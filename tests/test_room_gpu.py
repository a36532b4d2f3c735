"""Room sweep on the device (csrc/room.cu, p2pb_b200/room.py) against the CPU restatement of the reference's host code
(oracle/room.py <- /root/reference/denoise_room.py:352-421, 141-146, 262-289, 492-505): kernels bit-exact on indices / integer
sums, the whole sweep value-checked against an oracle-built sweep through the same network, and shard independence."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _room(n=24000, seed=0):
    """Noisy walls of a 3 x 2 x 1.5 m box (metres), fp32."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform([0, 0, 0], [3.0, 2.0, 1.5], size=(n, 3))
    face = rng.integers(0, 3, n)
    pts[np.arange(n), face] = np.where(rng.random(n) < 0.5, 0.0, np.array([3.0, 2.0, 1.5])[face])
    return (pts + rng.normal(0, 0.005, pts.shape)).astype(np.float32)


def _csr(room_t, centers_t, radius):
    from p2pb_b200 import ops

    off, csr = ops.radius_query(centers_t, room_t, radius)
    return off, csr, off.cpu().numpy(), csr.cpu().numpy().astype(np.int64)


def test_pad_fps_normalize_accumulate_kernels_match_the_oracle():
    from oracle import room as OR
    from p2pb_b200 import ops

    room = _room()
    room_t = torch.from_numpy(room).cuda()
    M, seed = 512, 42
    centers = room_t[torch.tensor([5, 1000, 7000, 12000, 20000], device="cuda")].contiguous()
    # radius 0.12: under-full patches, radius 0.5: over-full ones
    off, csr, off_h, csr_h = _csr(room_t, centers, 0.12)
    n = np.diff(off_h)
    assert (n < M).all() and (n > 8).all(), n
    jobs = torch.arange(5, dtype=torch.int32, device="cuda")
    keys = torch.tensor([100, 7, 3, 900, 12], dtype=torch.int32, device="cuda")          # global patch numbers != CSR-local ones
    xyz, idx, cut = ops.room_pad_patches(room_t, off, csr, jobs, M, seed, job_key=keys)
    for j in range(5):
        mp = csr_h[off_h[j]:off_h[j + 1]]
        ox, oi, oc = OR.pad_patch(room, mp, M, seed, int(keys[j]))
        assert int(cut[j]) == oc == len(mp)
        np.testing.assert_array_equal(idx[j].cpu().numpy(), oi)                           # same duplicates drawn (counter RNG, integer)
        np.testing.assert_array_equal(xyz[j, :oc].cpu().numpy(), ox[:oc])                 # the patch itself: exact copy
        sig = np.linalg.norm(room[mp].max(0) - room[mp].min(0)) * 1e-2
        np.testing.assert_allclose(xyz[j, oc:].cpu().numpy(), ox[oc:], rtol=0, atol=2e-5 * sig + 1e-7)   # jitter: libm vs CUDA log/cos
        jit = xyz[j, oc:].cpu().numpy() - room[oi[oc:]]
        assert 0.7 * sig < jit.std() < 1.3 * sig                                          # N(0, sigma^2) (denoise_room.py:377-378)
    # pre-drawn host randoms (--strict_ref): the kernel uses exactly what it is given
    d = M - n
    pre_off = torch.from_numpy(np.concatenate([[0], np.cumsum(d)])).cuda()
    rng = np.random.default_rng(1)
    pidx = np.concatenate([rng.integers(0, n[j], d[j]) for j in range(5)]).astype(np.int32)
    pnoise = rng.normal(0, 0.01, (int(d.sum()), 3)).astype(np.float32)
    xyz2, idx2, _ = ops.room_pad_patches(room_t, off, csr, jobs, M, seed, pre=(pre_off, torch.from_numpy(pidx).cuda(), torch.from_numpy(pnoise).cuda()))
    o = 0
    for j in range(5):
        mp = csr_h[off_h[j]:off_h[j + 1]]
        src = mp[pidx[o:o + d[j]]]
        np.testing.assert_array_equal(idx2[j, n[j]:].cpu().numpy(), src)
        np.testing.assert_array_equal(xyz2[j, n[j]:].cpu().numpy(), room[src] + pnoise[o:o + d[j]])
        o += d[j]
    # over-full patches: exact FPS from a given start, ragged n, bit-exact indices
    off, csr, off_h, csr_h = _csr(room_t, centers, 0.5)
    n = np.diff(off_h)
    assert (n >= M).all(), n
    job_patch = torch.tensor([0, 0, 1, 2, 3, 4, 4], dtype=torch.int32, device="cuda")
    starts_h = [OR.fps_start(seed, int(p), r, int(n[p])) for p, r in zip([0, 0, 1, 2, 3, 4, 4], [0, 1, 0, 0, 0, 0, 1])]
    xyz, idx = ops.room_fps_patches(room_t, off, csr, job_patch, torch.tensor(starts_h, dtype=torch.int32, device="cuda"), int(n.max()), M)
    for j, p in enumerate([0, 0, 1, 2, 3, 4, 4]):
        mp = csr_h[off_h[p]:off_h[p + 1]]
        ox, oi, _ = OR.fps_patch(room, mp, M, starts_h[j])
        np.testing.assert_array_equal(idx[j].cpu().numpy(), oi)
        np.testing.assert_array_equal(xyz[j].cpu().numpy(), ox)
    assert not torch.equal(idx[0], idx[1])                  # two replicas of one patch start elsewhere -> different subsets
    # normalisation (float64 statistics, denoise_room.py:141-146)
    x, c, s = ops.patch_normalize(xyz)
    ox, oc, os_ = OR.normalize(xyz.cpu().numpy())
    np.testing.assert_allclose(c.cpu().numpy(), oc[:, 0], rtol=0, atol=1e-12)
    np.testing.assert_allclose(s.cpu().numpy(), os_[:, 0, 0], rtol=1e-13)
    np.testing.assert_allclose(x.cpu().numpy(), ox, rtol=0, atol=1.2e-7)
    # accumulation: integer sums bit-exact vs the restatement, mean == the reference's sequential running mean
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = (x + 0.01 * torch.randn(x.shape, device="cuda", generator=g)).contiguous()
    cutv = torch.tensor([M, M, 300, M, 17, M, 0], dtype=torch.int32, device="cuda")
    sums = torch.zeros(room.shape[0], 3, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(room.shape[0], dtype=torch.int32, device="cuda")
    ops.room_accumulate(pred, c, s, idx, cutv, sums, cnt)
    s1, c1 = np.zeros((room.shape[0], 3), np.int64), np.zeros(room.shape[0], np.int32)
    OR.accumulate_fixed(pred.cpu().numpy(), c.cpu().numpy(), s.cpu().numpy(), idx.cpu().numpy(), cutv.cpu().numpy(), s1, c1)
    np.testing.assert_array_equal(sums.cpu().numpy(), s1)
    np.testing.assert_array_equal(cnt.cpu().numpy(), c1)
    from p2pb_b200.parallel import RoomAccumulator

    acc = RoomAccumulator(room.shape[0], "cuda")
    acc.sum, acc.count = sums, cnt
    world_pts = (pred.double() * s[:, None, None] + c[:, :, None]).permute(0, 2, 1).cpu().numpy()
    mean, num = OR.running_mean(room, world_pts, idx.cpu().numpy().astype(np.int64), cutv.cpu().numpy())
    np.testing.assert_allclose(acc.mean(room_t).cpu().numpy(), mean, rtol=0, atol=1e-10)


def _pvdl_small(tmp_path, npoints=1024, extra=0):
    import yaml as _yaml

    from p2pb_b200.config import load_yaml
    from p2pb_b200.model_loader import save_checkpoint, seeded_state_dict
    from p2pb_b200.p2pb import P2PB
    from p2pb_b200.unet_pvc import PVCNN2Unet

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = load_yaml(os.path.join(root, "p2pb_b200", "configs", "PVDL_SNPP.yaml")).to_dict()
    d["data"]["npoints"] = npoints
    d["model"]["extra_feature_channels"] = extra
    d["data"]["use_rgb_features"] = bool(extra)
    d["model"]["ema"] = False
    os.makedirs(tmp_path / "PVDL_test", exist_ok=True)
    (tmp_path / "PVDL_test" / "opt.yaml").write_text(_yaml.safe_dump(d))
    cfg = load_yaml(str(tmp_path / "PVDL_test" / "opt.yaml"))
    cfg.gpu = "cpu"
    net = PVCNN2Unet(cfg)
    net.load_state_dict(seeded_state_dict(net, 0, head_scale=0.02))
    save_checkpoint(str(tmp_path / "PVDL_test" / "step_100.pth"), P2PB(cfg, net), step=100)
    cfg.gpu = "cuda:0"
    return cfg, P2PB(cfg, net.cuda()).eval()


def test_sweep_matches_oracle_built_sweep_and_is_shard_independent(tmp_path):
    """The whole sweep (plan -> device patch creation -> normalise -> sample -> fixed-point accumulate -> mean) equals a sweep
    whose patches, normalisation and running mean are built by the CPU restatement of the reference's host code around the SAME
    network; and two half-shards add up to the single-rank accumulators bit for bit."""
    from oracle import room as OR
    from p2pb_b200 import ops
    from p2pb_b200 import room as R

    M, k, radius, steps, bs, seed = 1024, 2, 0.5, 2, 8, 42
    cfg, model = _pvdl_small(tmp_path, npoints=M, extra=3)
    room = _room(40000, seed=3)             # ~1200 points per radius-0.5 disc on a wall: a mix of under- and over-full patches
    rng = np.random.default_rng(5)
    rgb = rng.random((room.shape[0], 3)).astype(np.float32)
    room_t, feats_t = torch.from_numpy(room).cuda(), torch.from_numpy(rgb).cuda()
    res = R.sweep(model, room_t, M, k, radius, steps, bs, seed, feats=feats_t)
    n_centers = int(np.ceil(room.shape[0] / M) * k)
    assert res.n_jobs == res.n_jobs_rank and res.n_jobs > n_centers        # some patches are over-full (several FPS replicas)
    # ---- oracle-built sweep: same plan (centres by the oracle FPS, radius sets by brute force), CPU patch creation
    from oracle import ops as OO

    cidx = OO.furthest_point_sampling_forward(torch.from_numpy(room.T.copy())[None], n_centers)[0].long().numpy()
    xyz_l, idx_l, cut_l = [], [], []
    o_off, o_idx = OO.radius_query(torch.from_numpy(room[cidx]), torch.from_numpy(room), radius)      # oracle radius sets
    o_off, o_idx = o_off.numpy(), o_idx.numpy().astype(np.int64)
    for p in range(n_centers):
        mp = o_idx[o_off[p]:o_off[p + 1]]
        if len(mp) == 0:
            continue
        if len(mp) < M:
            x, i, c = OR.pad_patch(room, mp, M, seed, p)
            xyz_l.append(x), idx_l.append(i), cut_l.append(c)
        else:
            for r in range(len(mp) // M + 1):
                x, i, _ = OR.fps_patch(room, mp, M, OR.fps_start(seed, p, r, len(mp)))
                xyz_l.append(x), idx_l.append(i), cut_l.append(M)
    assert len(xyz_l) == res.n_jobs and 0 < sum(c < M for c in cut_l) < len(cut_l)
    xyz_o, idx_o, cut_o = np.stack(xyz_l), np.stack(idx_l), np.array(cut_l)
    xs, center, scale = OR.normalize(xyz_o)
    world = []
    for s in range(0, len(xs), bs):
        xb = torch.from_numpy(xs[s:s + bs]).cuda()
        n_b = xb.shape[0]
        if n_b < bs:
            xb = torch.cat([xb, xb[-1:].expand(bs - n_b, -1, -1)]).contiguous()
        ib = torch.from_numpy(idx_o[s:s + bs]).cuda()
        if n_b < bs:
            ib = torch.cat([ib, ib[-1:].expand(bs - n_b, -1)])
        cond = feats_t[ib.reshape(-1)].reshape(bs, M, 3).permute(0, 2, 1).contiguous()
        out = model.sample(x_start=xb, x_cond=cond, verbose=False, steps=steps, use_ema=False, log_count=1)["x_pred"][:n_b]
        world.append(out.permute(0, 2, 1).double().cpu().numpy() * scale[s:s + n_b] + center[s:s + n_b])     # denoise_room.py:176
    mean, num = OR.running_mean(room, np.concatenate(world), idx_o, cut_o)
    np.testing.assert_array_equal(res.count.cpu().numpy(), num.astype(np.int64))
    err = np.abs(res.denoised.cpu().numpy() - mean)
    moved = np.abs(mean - room)[num > 0]
    print(f"sweep vs oracle-built sweep: max|err|={err.max():.3e} mean|err|={err.mean():.3e} p50={np.percentile(err, 50):.1e} "
          f"p99={np.percentile(err, 99):.1e}; |moved| mean={moved.mean():.3e} max={moved.max():.3e}; "
          f"{int((num > 0).sum())}/{len(num)} points updated, {res.n_jobs} jobs")
    # identical patches / indices / counts; the inputs of the network differ by <= 1 fp32 ulp (normalisation) and 1e-5 sigma
    # (libm vs CUDA log/cos in the padding jitter), which a voxel-rounding flip inside the network can amplify on a few points:
    # median at the f64 rounding level, mean three orders below the displacement, worst point well below it
    assert np.percentile(err, 50) <= 1e-9 and err.mean() <= 2e-3 * moved.mean() and err.max() <= 0.25 * moved.max()
    assert moved.mean() > 1e-4                                      # the network did move the points
    # ---- shard independence: ranks (0,2) + (1,2) == rank (0,1), integers
    a0, _, _, J, n0 = R.sweep_shard(model, room_t, M, k, radius, steps, bs, seed, feats=feats_t, rank=0, world=2)
    a1, _, _, _, n1 = R.sweep_shard(model, room_t, M, k, radius, steps, bs, seed, feats=feats_t, rank=1, world=2)
    a, _, _, _, _ = R.sweep_shard(model, room_t, M, k, radius, steps, bs, seed, feats=feats_t)
    assert n0 + n1 == J and abs(n0 - n1) <= 1
    assert torch.equal(a0.count + a1.count, a.count)
    d = (a0.sum + a1.sum - a.sum).abs().max().item()
    print(f"two half-shards vs one rank: max |sum difference| = {d} fixed-point units (2^-40)")
    assert d <= 2 ** 18        # the network runs in different batch compositions -> fp32-rounding-level differences only (2^-22 m)


def test_denoise_room_cli_flags(tmp_path):
    """denoise_room.py end to end: default output name (denoise_room.py:430-445), --intermediate step files, the not-updated fill,
    --average_predictions False, --strict_ref; values checked against the sweep API."""
    import denoise_room as D
    from p2pb_b200.io_ply import read_ply, write_ply

    cfg, model = _pvdl_small(tmp_path, npoints=1024, extra=0)
    room = _room(16000, seed=9)
    (tmp_path / "scene" / "scans").mkdir(parents=True)
    rp = tmp_path / "scene" / "scans" / "room_a.ply"
    write_ply(str(rp), room.astype(np.float64), None)
    ck = str(tmp_path / "PVDL_test" / "step_100.pth")
    common = ["--room_path", str(rp), "--model_path", ck, "--steps", "2", "--batch_size", "8", "--k", "2", "--use_ema", ""]
    D.main(common + ["--intermediate"])
    out_path = tmp_path / "scene" / "predictions" / "P2SB" / "PVDL-test_room-a_100_2.ply"
    assert out_path.exists(), os.listdir(tmp_path / "scene")
    out, _ = read_ply(str(out_path))
    assert out.shape == room.shape and np.isfinite(out).all()
    steps = [read_ply(os.path.splitext(str(out_path))[0] + f"_step_{i}.ply")[0] for i in range(2)]
    # step 0 of the chain is the final state (denoise_room.py:160-163): equal on every point a patch updated; the points no
    # patch touched copy a random other point, drawn anew per file (denoise_room.py:540-550)
    upd = (steps[0] == out).all(axis=1)
    assert 0.3 < upd.mean() < 1.0
    assert np.abs(steps[1][upd] - out[upd]).max() > 1e-6
    moved = np.linalg.norm(out - room, axis=1)
    assert 1e-4 < moved[upd].mean() < 0.05 and (moved[upd] < 0.5).mean() > 0.999      # (a filled point may coincide in both files)
    assert moved[~upd].max() > 0.5                                  # filled points sit somewhere else in the room
    # a second call finds the prediction and returns without touching it (denoise_room.py:447-449)
    mt = os.path.getmtime(out_path)
    D.main(common)
    assert os.path.getmtime(out_path) == mt
    # strict_ref drops one patch per chunk: fewer contributions, still every flag path runs
    o2 = tmp_path / "strict.ply"
    D.main(common + ["--out_path", str(o2), "--strict_ref"])
    s2, _ = read_ply(str(o2))
    assert s2.shape == room.shape and np.abs(s2 - out).max() > 0
    # no averaging: FPS of all denoised patches down to the room size
    o3 = tmp_path / "loose.ply"
    D.main(common + ["--out_path", str(o3), "--average_predictions", "False", "--k", "4"])
    s3, _ = read_ply(str(o3))
    assert s3.shape == room.shape and len(np.unique(s3, axis=0)) > 0.9 * len(s3)
    from p2pb_b200 import ops

    cd = ops.calculate_cd(torch.from_numpy(s3).float().cuda()[None], torch.from_numpy(room).cuda()[None])
    assert cd[0] < 1e-2, cd          # point spacing of this room is ~3.7 cm (squared 1.4e-3, counted both ways)
