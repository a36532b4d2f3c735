"""Data-parallel sharding of patch batches over ranks and the single exchange of the path: room reassembly.

Patches are independent (GroupNorm / SE / attention are per sample), so the job list is split contiguously over the
ranks with NO data-path collective; each rank creates and denoises only its shard with its own replica of the weights.  The
only exchange is at the end of ``denoise_room``: the reference keeps a sequential per-point running mean over the un-padded
points of every patch (``denoise_room.py:262-289``, numba, float64); here every rank accumulates per-point FIXED-POINT sums
(int64, units of 2^-40) and counts with integer atomics (``csrc/room.cu``) and the ranks ``all_reduce(SUM)`` both -- NCCL
over NVLink/NVSwitch on the GPUs (28 B/point: 28-140 MB for 1-5 M-point rooms), gloo in the CPU tests.  Integer addition is
associative, so the reassembled room is bit-identical for any patch order, batch split and rank count; it equals the
reference's running mean up to f64 rounding (|x| 2^-40 per contribution).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

FIXED_ONE = float(2 ** 40)


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first ``n_items % world`` ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class RoomAccumulator:
    """Per-point sums (int64 fixed point) and update counts of one room on one rank."""

    def __init__(self, n_points: int, device="cuda"):
        self.sum = torch.zeros(n_points, 3, dtype=torch.int64, device=device)
        self.count = torch.zeros(n_points, dtype=torch.int32, device=device)

    def add(self, x_pred: torch.Tensor, center: torch.Tensor, scale: torch.Tensor, idx: torch.Tensor, cut: torch.Tensor) -> None:
        """Network output of P patches ``x_pred [P,3,M]`` (normalised) + their centre / scale / room indices / cut."""
        from . import ops

        ops.room_accumulate(x_pred.contiguous(), center, scale, idx, cut, self.sum, self.count)

    def reduce(self) -> None:
        """all_reduce(SUM) over the ranks (if a process group is up)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.sum, op=dist.ReduceOp.SUM)
            dist.all_reduce(self.count, op=dist.ReduceOp.SUM)

    def mean(self, room: torch.Tensor) -> torch.Tensor:
        """f64 [N,3]: sum / count where a point was updated, the input point elsewhere (denoise_room.py:466-468 initialises
        the prediction with the noisy room)."""
        upd = self.count > 0
        m = self.sum.to(torch.float64) / FIXED_ONE / self.count.clamp(min=1).to(torch.float64).unsqueeze(1)
        return torch.where(upd.unsqueeze(1), m, room.to(torch.float64))
