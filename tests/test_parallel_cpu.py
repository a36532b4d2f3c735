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


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from p2pb_b200 import parallel as PP

    n_patches, n_room, npts = 11, 500, 64
    lo, hi = PP.shard_range(n_patches, rank, world)
    g = torch.Generator().manual_seed(0)
    idx = torch.stack([torch.randperm(n_room, generator=g)[:npts] for _ in range(n_patches)])      # same on every rank
    pts = torch.randn(n_patches, npts, 3, generator=g, dtype=torch.float32)
    cut = torch.randint(npts // 2, npts + 1, (n_patches,), generator=g)
    acc = PP.RoomAccumulator(n_room, device="cpu")
    for p in range(lo, hi):
        acc.add(idx[p, : cut[p]], pts[p, : cut[p]])
    mean, count = acc.reduce()            # all_reduce(SUM) of the f64 sums and int counts, then divide
    q.put((rank, lo, hi, mean.numpy(), count.numpy()))
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
    # sequential reference (the reference's running mean, float64)
    n_patches, n_room, npts = 11, 500, 64
    g = torch.Generator().manual_seed(0)
    idx = torch.stack([torch.randperm(n_room, generator=g)[:npts] for _ in range(n_patches)])
    pts = torch.randn(n_patches, npts, 3, generator=g, dtype=torch.float32)
    cut = torch.randint(npts // 2, npts + 1, (n_patches,), generator=g)
    mean = np.zeros((n_room, 3)); cnt = np.zeros(n_room, dtype=np.int64)
    for p in range(n_patches):
        for j in range(int(cut[p])):
            i = int(idx[p, j]); cnt[i] += 1
            mean[i] += (pts[p, j].double().numpy() - mean[i]) / cnt[i]
    np.testing.assert_array_equal(res[0][4], cnt)
    np.testing.assert_allclose(res[0][3], mean, rtol=0, atol=1e-12)
