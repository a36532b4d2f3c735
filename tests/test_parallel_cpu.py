"""CPU tests of the multi-GPU host logic (N>1 path) with a world_size-2 gloo group: patch sharding and the room
reassembly exchange (the only collective of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    n_patches, n_room, npts = 11, 500, 64
    g = torch.Generator().manual_seed(0)
    idx = torch.stack([torch.randperm(n_room, generator=g)[:npts] for _ in range(n_patches)]).to(torch.int32)   # same on every rank
    x_pred = torch.randn(n_patches, 3, npts, generator=g, dtype=torch.float32) * 0.5            # normalised network output
    center = torch.randn(n_patches, 3, generator=g, dtype=torch.float64) * 3.0
    scale = torch.rand(n_patches, generator=g, dtype=torch.float64) + 0.25
    cut = torch.randint(npts // 2, npts + 1, (n_patches,), generator=g).to(torch.int32)
    room = torch.randn(n_room, 3, generator=g, dtype=torch.float32)
    return n_patches, n_room, npts, idx, x_pred, center, scale, cut, room


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import room as OR
    from p2pb_b200 import parallel as PP

    n_patches, n_room, npts, idx, x_pred, center, scale, cut, room = _case()
    lo, hi = PP.shard_range(n_patches, rank, world)
    acc = PP.RoomAccumulator(n_room, device="cpu")
    # the accumulate kernel is CUDA-only (tests/test_room_gpu.py pins it to this restatement bit for bit): fill this rank's
    # shard through the checker, then exercise the product's exchange + mean
    OR.accumulate_fixed(x_pred[lo:hi].numpy(), center[lo:hi].numpy(), scale[lo:hi].numpy(), idx[lo:hi].numpy(), cut[lo:hi].numpy(),
                        acc.sum.numpy(), acc.count.numpy())
    acc.reduce()            # all_reduce(SUM) of the int64 fixed-point sums and the int32 counts
    q.put((rank, lo, hi, acc.mean(room).numpy(), acc.count.numpy(), acc.sum.numpy()))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from p2pb_b200 import parallel as PP

    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 4, 8):
            r = [PP.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_room_reassembly_two_ranks_matches_sequential_running_mean():
    """denoise_room.py:262-289 keeps a sequential per-point running mean; sharded sums + all_reduce + divide must agree."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 11
    np.testing.assert_array_equal(res[0][3], res[1][3])       # every rank ends with the same reassembled room
    # single-rank accumulation of all patches gives the SAME integers (order / shard independence of the fixed-point sums)
    from oracle import room as OR

    n_patches, n_room, npts, idx, x_pred, center, scale, cut, room = _case()
    s1, c1 = np.zeros((n_room, 3), np.int64), np.zeros(n_room, np.int32)
    order = np.random.default_rng(0).permutation(n_patches)
    OR.accumulate_fixed(x_pred.numpy()[order], center.numpy()[order], scale.numpy()[order], idx.numpy()[order], cut.numpy()[order], s1, c1)
    np.testing.assert_array_equal(res[0][5], s1)
    np.testing.assert_array_equal(res[0][4], c1)
    # the reference's sequential running mean (update_prediction_noisy_batches, float64) on the de-normalised patches
    world_pts = (x_pred.double() * scale[:, None, None] + center[:, :, None]).permute(0, 2, 1).numpy()
    mean, num = OR.running_mean(room.numpy(), world_pts, idx.numpy().astype(np.int64), cut.numpy())
    np.testing.assert_array_equal(num.astype(np.int64), res[0][4])
    np.testing.assert_allclose(res[0][3], mean, rtol=0, atol=1e-10)
    assert (num == 0).any() and np.array_equal(res[0][3][num == 0], room.numpy()[num == 0].astype(np.float64))   # untouched points keep the input


def test_reference_chunking_quirk_and_counter_rng_match_the_oracle():
    """--strict_ref drops exactly the patches the reference never denoises (np.array_split chunks, `[start:end]` with end = last
    index, denoise_room.py:492-505); the FPS start draws of the product equal the oracle's restatement of the counter RNG."""
    from oracle import room as OR
    from p2pb_b200 import room as R

    for n, bs in ((1, 32), (31, 32), (32, 32), (33, 32), (100, 32), (257, 8), (1000, 32)):
        np.testing.assert_array_equal(R.reference_kept_jobs(n, bs), OR.reference_kept_patches(n, bs))
    for seed, patch, rep, n in ((42, 0, 0, 9000), (42, 17, 2, 12345), (7, 979, 3, 30000)):
        assert R.fps_start(seed, patch, rep, n) == OR.fps_start(seed, patch, rep, n)
