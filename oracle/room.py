"""CPU restatement of the room sweep around the hot path -- patch creation, per-patch normalisation, reassembly -- TEST
INFRASTRUCTURE ONLY.  Follows /root/reference/denoise_room.py:

  create_patches                   :352-421   (under-full: pad with jittered duplicates; over-full: `fraction` FPS subsets)
  denoise_patch_batch              :141-146, 176  (centre / max-norm scale in float64, de-normalise)
  update_prediction_noisy_batches  :262-289   (sequential running mean, first update replaces)
  main                             :492-505   (np.array_split chunks, `[start:end]` with end = last index drops one patch per chunk)
                                   :540-550   (points never updated copy a random other point)

Where the reference draws from np.random / fpsample (un-vendored; start index random) the product uses a counter-based RNG
keyed by (seed, patch, slot) so that the result is independent of rank count and launch shape; ``rm_draw`` restates it
bit for bit (splitmix64 finaliser, csrc/room.cu).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import ops as OO

_M64 = (1 << 64) - 1


def _mix(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def rm_draw(seed: int, patch: int, slot: int, draw: int) -> int:
    return _mix((_mix((seed ^ ((patch << 32) | slot)) & _M64) + draw) & _M64)


def _uniform(h: int) -> np.float32:
    return np.float32((np.float32(h >> 40) + np.float32(0.5)) * np.float32(1.0 / 16777216.0))


def pad_patch(room: np.ndarray, mapping: np.ndarray, M: int, seed: int, patch: int):
    """denoise_room.py:369-395 with the counter-based RNG -> (xyz [M,3] f32, idx [M], cut)."""
    pts = room[mapping].astype(np.float32)
    n = len(pts)
    sigma = np.float32(np.sqrt(((pts.max(0).astype(np.float64) - pts.min(0).astype(np.float64)) ** 2).sum()) * 1e-2)
    xyz = np.empty((M, 3), np.float32)
    idx = np.empty((M,), np.int64)
    xyz[:n], idx[:n] = pts, mapping
    two = np.float32(2.0)
    for s in range(n, M):
        loc = rm_draw(seed, patch, s, 0) % n
        u1, u2, u3, u4 = (_uniform(rm_draw(seed, patch, s, d)) for d in (1, 2, 3, 4))
        r1 = np.sqrt(np.float32(-2.0) * np.log(u1), dtype=np.float32)
        r2 = np.sqrt(np.float32(-2.0) * np.log(u3), dtype=np.float32)
        noise = np.array([sigma * r1 * np.cos(np.float32(np.pi) * two * u2, dtype=np.float32),
                          sigma * r1 * np.sin(np.float32(np.pi) * two * u2, dtype=np.float32),
                          sigma * r2 * np.cos(np.float32(np.pi) * two * u4, dtype=np.float32)], np.float32)
        xyz[s] = pts[loc] + noise
        idx[s] = mapping[loc]
    return xyz, idx, n


def fps_start(seed: int, patch: int, replica: int, n: int) -> int:
    """Start index of replica `replica` of an over-full patch (draw 5 of slot `replica`)."""
    return rm_draw(seed, patch, replica, 5) % n


def fps_patch(room: np.ndarray, mapping: np.ndarray, M: int, start: int):
    """denoise_room.py:404-411: exact FPS of M of the patch's points from `start` (reference FPS tie-break)."""
    import torch

    pts = np.ascontiguousarray(room[mapping].astype(np.float32).T)        # [3, n]
    n = pts.shape[1]
    sel = np.zeros((M,), np.int32)
    OO.lib().ora_fps_from_start(ctypes.c_void_p(pts.ctypes.data), int(n), int(M), int(start), ctypes.c_void_p(sel.ctypes.data))
    return room[mapping][sel].astype(np.float32), mapping[sel], sel


def normalize(xyz: np.ndarray):
    """denoise_room.py:141-146 (float64): xyz [P,M,3] -> x_start [P,3,M] f32, center [P,1,3] f64, scale [P,1,1] f64."""
    p = xyz.astype(np.float64)
    center = p.mean(axis=1, keepdims=True)
    p = p - center
    scale = np.linalg.norm(p, axis=2, keepdims=True).max(axis=1, keepdims=True)
    return np.ascontiguousarray((p / scale).astype(np.float32).transpose(0, 2, 1)), center, scale


def running_mean(room: np.ndarray, patches: np.ndarray, idxs: np.ndarray, cuts: np.ndarray):
    """update_prediction_noisy_batches (denoise_room.py:262-289), sequential, float64 -> (denoised [N,3], num_updates [N])."""
    denoised = room.astype(np.float64).copy()
    num = np.zeros(room.shape[0])
    for patch, idx, cut in zip(patches, idxs, cuts):
        patch, idx = patch[:cut].astype(np.float64), idx[:cut]
        num[idx] += 1
        first = (num[idx] == 1)[:, None]
        denoised[idx] = np.where(first, patch, (denoised[idx] * (num[idx] - 1)[:, None] + patch) / num[idx][:, None])
    return denoised, num


def reference_kept_patches(n_patches: int, batch_size: int) -> np.ndarray:
    """denoise_room.py:492-505: chunks = np.array_split(arange(P), ceil(P / batch_size)); each chunk is processed as
    [chunk[0] : chunk[-1]] -- the last patch of every chunk is never denoised."""
    n_batches = int(np.ceil(n_patches / batch_size))
    keep = []
    for ch in np.array_split(np.arange(n_patches), n_batches):
        keep.extend(range(ch[0], ch[-1]))
    return np.array(keep, dtype=np.int64)


FIXED_ONE = float(2 ** 40)


def accumulate_fixed(x_pred: np.ndarray, center: np.ndarray, scale: np.ndarray, idx: np.ndarray, cut: np.ndarray,
                     sum_fixed: np.ndarray, count: np.ndarray) -> None:
    """Restatement of room_accumulate_kernel (csrc/room.cu): de-normalise in float64 (denoise_room.py:176), round to 2^-40 units,
    integer adds.  x_pred [P,3,M], center [P,3], scale [P], idx [P,M], cut [P]; sum_fixed int64 [N,3], count int32 [N]."""
    for p in range(x_pred.shape[0]):
        c = int(cut[p])
        v = x_pred[p, :, :c].astype(np.float64).T * float(scale[p]) + center[p].astype(np.float64)[None, :]
        np.add.at(sum_fixed, idx[p, :c], np.rint(v * FIXED_ONE).astype(np.int64))
        np.add.at(count, idx[p, :c], 1)
