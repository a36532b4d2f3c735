"""Room sweep around the hot path (``denoise_room.py:424-570``), device-side: patch planning, patch creation, batched sampling,
reassembly.  Host Python only PLANS (job lists, shards); the data stays in HBM from the room upload to the reassembled room:

  FPS centres (ops.furthest_point_sampling) -> radius counts -> job list (under-full patch -> 1 padded job, over-full patch ->
  n // npoints + 1 FPS replicas, :396-419) -> contiguous shard per rank -> radius CSR of the shard's patches only ->
  room_pad_patches / room_fps_patches (one launch each) -> per batch: patch_normalize -> P2PB.sample -> room_accumulate ->
  ONE all_reduce -> mean.

Randomness: the reference draws padding duplicates / jitter from np.random and FPS start points inside ``fpsample`` (un-vendored).
Here both come from a counter-based RNG keyed by (seed, patch, slot), so the jobs -- and therefore the denoised room -- do not
depend on the number of ranks.  ``strict_ref`` reproduces two reference behaviours instead: padding randoms are drawn on the
host from np.random in the reference's order, and the last patch of every ``np.array_split`` chunk is dropped (:492-505).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import ops
from .parallel import RoomAccumulator, shard_range

_M64 = (1 << 64) - 1


def _mix(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def fps_start(seed: int, patch: int, replica: int, n: int) -> int:
    """Start index of FPS replica ``replica`` of over-full patch ``patch`` (draw 5 of the counter RNG in csrc/room.cu)."""
    return _mix((_mix((seed ^ ((patch << 32) | replica)) & _M64) + 5) & _M64) % n


def reference_kept_jobs(n_total: int, batch_size: int) -> np.ndarray:
    """denoise_room.py:492-505: the reference splits the patch list with np.array_split into ceil(P / batch_size) chunks and
    processes ``[chunk[0] : chunk[-1]]`` -- the last patch of every chunk is never denoised.  -> indices it does denoise."""
    nb = int(np.ceil(n_total / batch_size))
    return np.concatenate([ch[:-1] for ch in np.array_split(np.arange(n_total), nb)]).astype(np.int64)


@dataclass
class Plan:
    centers: torch.Tensor          # [P,3] device
    counts: np.ndarray             # [P] points within the radius of every centre
    job_patch: np.ndarray          # [J] patch of every job, reference order (patch-major, replicas consecutive)
    job_replica: np.ndarray        # [J] replica number (0 for padded jobs)
    n_total: int                   # J before strict_ref dropping


def plan_jobs(room: torch.Tensor, npoints: int, k: int, radius: float, batch_size: int, strict_ref: bool = False) -> Plan:
    """denoise_room.py:447-465 + the job structure of create_patches (:352-421)."""
    N = room.shape[0]
    n_centers = int(np.ceil(N / npoints) * k)
    center_idx = ops.furthest_point_sampling(room.t().contiguous().unsqueeze(0), n_centers)[0].long()
    centers = room[center_idx].contiguous()
    counts = ops.radius_count(centers, room, radius).cpu().numpy().astype(np.int64)
    reps = np.where(counts == 0, 0, np.where(counts < npoints, 1, counts // npoints + 1))
    job_patch = np.repeat(np.arange(n_centers), reps)
    job_replica = np.concatenate([np.arange(r) for r in reps]) if len(reps) else np.zeros(0, np.int64)
    n_total = len(job_patch)
    if strict_ref and n_total:
        keep = reference_kept_jobs(n_total, batch_size)
        job_patch, job_replica = job_patch[keep], job_replica[keep]
    return Plan(centers, counts, job_patch.astype(np.int64), job_replica.astype(np.int64), n_total)


@dataclass
class Patches:
    xyz: torch.Tensor              # [J,M,3] world coordinates
    idx: torch.Tensor              # int32 [J,M] room indices
    cut: torch.Tensor              # int32 [J] rows that update the room


def create_patches(room: torch.Tensor, plan: Plan, lo: int, hi: int, npoints: int, radius: float, seed: int,
                   strict_ref: bool = False) -> Patches:
    """Jobs [lo, hi) of the plan -> device tensors, in job order."""
    dev = room.device
    jp, jr = plan.job_patch[lo:hi], plan.job_replica[lo:hi]
    J = len(jp)
    if J == 0:
        return Patches(torch.empty((0, npoints, 3), device=dev), torch.empty((0, npoints), dtype=torch.int32, device=dev),
                       torch.empty((0,), dtype=torch.int32, device=dev))
    uniq, local = np.unique(jp, return_inverse=True)              # radius CSR of this shard's patches only
    off, csr = ops.radius_query(plan.centers[torch.from_numpy(uniq).to(dev)].contiguous(), room, radius)
    n = plan.counts[jp]
    small = n < npoints
    xyz = torch.empty((J, npoints, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((J, npoints), dtype=torch.int32, device=dev)
    cut = torch.full((J,), npoints, dtype=torch.int32, device=dev)
    to_dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
    if small.any():
        js = np.nonzero(small)[0]
        pre = None
        if strict_ref:           # the reference's np.random sequence (denoise_room.py:371-378): randint, then normal, per patch in order
            offs, pidx, pnoise = [0], [], []
            off_h = off.cpu().numpy()
            room_h = None
            for j in js:
                nn, d = int(n[j]), npoints - int(n[j])
                pidx.append(np.random.randint(0, nn, d))
                if room_h is None:
                    room_h, csr_h = room.cpu().numpy(), csr.cpu().numpy()
                pts = room_h[csr_h[off_h[local[j]]:off_h[local[j] + 1]]]
                sigma = np.linalg.norm(pts.max(0) - pts.min(0)) * 1e-2
                pnoise.append(np.random.normal(0, sigma, (d, 3)))
                offs.append(offs[-1] + d)
            pre = (to_dev(np.array(offs), torch.int64), to_dev(np.concatenate(pidx), torch.int32),
                   to_dev(np.concatenate(pnoise), torch.float32))
        # the counter RNG is keyed by the GLOBAL patch number (job_key), the CSR is addressed by the shard-local one
        x, i, c = ops.room_pad_patches(room, off, csr, to_dev(local[js], torch.int32), npoints, seed, pre,
                                       job_key=to_dev(jp[js], torch.int32))
        sel = to_dev(js, torch.int64)
        xyz[sel], idx[sel], cut[sel] = x, i, c
    if (~small).any():
        jl = np.nonzero(~small)[0]
        starts = np.array([fps_start(seed, int(jp[j]), int(jr[j]), int(n[j])) for j in jl], dtype=np.int32)
        x, i = ops.room_fps_patches(room, off, csr, to_dev(local[jl], torch.int32), to_dev(starts, torch.int32), int(n[jl].max()), npoints)
        sel = to_dev(jl, torch.int64)
        xyz[sel], idx[sel] = x, i
    return Patches(xyz, idx, cut)


def balanced_batch(n_jobs: int, batch_size: int) -> int:
    """Largest-needed batch for ceil(n_jobs / batch_size) equally filled sample() calls."""
    if n_jobs <= 0:
        return batch_size
    nb = -(-n_jobs // batch_size)
    return -(-n_jobs // nb)


@dataclass
class SweepResult:
    denoised: Optional[torch.Tensor]        # f64 [N,3] (average_predictions) or fp32 [N,3] (FPS of all denoised patches); rank 0
    count: Optional[torch.Tensor]           # int32 [N] updates per point (average_predictions)
    steps: Optional[List[torch.Tensor]]     # per logged step: f64 [N,3] (--intermediate)
    n_jobs: int
    n_jobs_rank: int


@torch.no_grad()
def sweep_shard(model, room: torch.Tensor, npoints: int, k: int, radius: float, steps: int, batch_size: int, seed: int,
                feats: Optional[torch.Tensor] = None, use_ema: bool = False, average_predictions: bool = True,
                intermediate: bool = False, strict_ref: bool = False, rank: int = 0, world: int = 1):
    """This rank's part of the sweep, before the exchange -> (acc, acc_steps, loose points, n_jobs, n_jobs_rank)."""
    dev = room.device
    N = room.shape[0]
    plan = plan_jobs(room, npoints, k, radius, batch_size, strict_ref)
    J = len(plan.job_patch)
    lo, hi = shard_range(J, rank, world)
    pt = create_patches(room, plan, lo, hi, npoints, radius, seed, strict_ref)
    acc = RoomAccumulator(N, dev) if average_predictions else None
    acc_steps = [RoomAccumulator(N, dev) for _ in range(steps)] if (intermediate and average_predictions) else None
    loose = []
    nj = hi - lo
    # balanced batches: the same number of sample() calls as ceil(nj / batch_size), all (nearly) full -- a shard of 261 jobs runs
    # as 9 x 29 instead of 8 x 32 + 5 padded to 32 (the captured graph has a static batch shape; one shape per sweep)
    batch_size = balanced_batch(nj, batch_size)
    for s in range(0, nj, batch_size):
        e = min(s + batch_size, nj)
        take = torch.arange(s, e, device=dev)
        if e - s < batch_size:                          # static batch shape for the captured graph: repeat the last job, cut = 0
            take = torch.cat([take, take[-1:].expand(batch_size - (e - s))])
        xyz, idx, cut = pt.xyz[take].contiguous(), pt.idx[take].contiguous(), pt.cut[take].clone()
        cut[e - s:] = 0
        x, center, scale = ops.patch_normalize(xyz)
        cond = None
        if feats is not None:
            cond = feats[idx.long().reshape(-1)].reshape(batch_size, npoints, -1).permute(0, 2, 1).contiguous()
        out = model.sample(x_start=x, x_cond=cond, verbose=False, steps=steps, use_ema=use_ema, log_count=steps if intermediate else 1)
        if average_predictions:
            acc.add(out["x_pred"], center, scale, idx, cut)
            if acc_steps is not None:
                for i in range(steps):                  # x_chain [B, T, 3, M], index 0 = final state (denoise_room.py:160-163)
                    acc_steps[i].add(out["x_chain"][:, i], center, scale, idx, cut)
        else:
            # denoise_room.py:523-531: FPS-order every denoised patch (world coordinates), keep them all
            w = (out["x_pred"][: e - s].to(torch.float64) * scale[: e - s, None, None] + center[: e - s, :, None]).float().contiguous()
            sel = ops.furthest_point_sampling(w, npoints)
            loose.append(torch.gather(w, 2, sel.long().unsqueeze(1).expand(-1, 3, -1)).transpose(1, 2).reshape(-1, 3))
    return acc, acc_steps, loose, J, nj


@torch.no_grad()
def sweep(model, room: torch.Tensor, npoints: int, k: int, radius: float, steps: int, batch_size: int, seed: int,
          feats: Optional[torch.Tensor] = None, use_ema: bool = False, average_predictions: bool = True, intermediate: bool = False,
          strict_ref: bool = False, rank: int = 0, world: int = 1) -> SweepResult:
    """The room sweep of ``denoise_room.py::main`` (:447-556) for ``room [N,3]`` (fp32, device) and optional per-point
    conditioning ``feats [N,F]``: every rank plans all jobs (cheap, deterministic), creates and denoises its contiguous shard,
    then one all_reduce."""
    import torch.distributed as dist

    dev, N = room.device, room.shape[0]
    acc, acc_steps, loose, J, nj = sweep_shard(model, room, npoints, k, radius, steps, batch_size, seed, feats, use_ema,
                                               average_predictions, intermediate, strict_ref, rank, world)
    if average_predictions:
        acc.reduce()
        den = acc.mean(room)
        stp = None
        if acc_steps is not None:
            stp = []
            for a in acc_steps:
                a.reduce()
                stp.append(a.mean(room))
        return SweepResult(den, acc.count, stp, J, nj)
    # not averaging (denoise_room.py:552-556): gather every rank's denoised patch points, FPS down to the room size on rank 0
    mine = torch.cat(loose, 0) if loose else torch.empty((0, 3), device=dev)
    if world > 1:
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.shape[0]], dtype=torch.int64, device=dev))
        mx = int(max(int(s.item()) for s in sizes))
        buf = torch.zeros((mx, 3), device=dev)
        buf[: mine.shape[0]] = mine
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        mine = torch.cat([p[: int(s.item())] for p, s in zip(parts, sizes)], 0)
    den = None
    if rank == 0:
        if mine.shape[0] < N:
            raise RuntimeError(f"average_predictions=False needs at least as many denoised patch points ({mine.shape[0]}) as room points ({N})")
        sel = ops.furthest_point_sampling(mine.t().contiguous().unsqueeze(0), N)[0].long()
        den = mine[sel]
    return SweepResult(den, None, None, J, nj)


def fill_not_updated(denoised: np.ndarray, count: np.ndarray) -> int:
    """denoise_room.py:540-550: points no patch touched copy the (denoised) position of a random point (np.random)."""
    missing = np.where(count == 0)[0]
    if len(missing):
        denoised[missing] = denoised[np.random.choice(len(denoised), len(missing))]
    return len(missing)
