"""Host-side plumbing of the batch-sharded (data-parallel) hot path -- SURVEY.md 8e.

Every kernel on the path is independent across the batch index (the reference's kernels take ``blockIdx.x = batch``,
selective_scan_fwd_kernel.cuh:80; ``program_id(2) = i_b``, csm_triton.py:20; the STFT is per clip), so N ranks each run
the path on a contiguous shard of the clips and there is no exchange step inside it.  The only collective is the one a
data-parallel trainer issues anyway: the SUM (then mean) of the parameter gradients, of which the scan contributes
``dA, dD, ddelta_bias``.  One process per GPU; ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is plumbing.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard ``[lo, hi)`` of ``n_clips`` for ``rank`` (the first ``n_clips % world`` ranks get
    one clip more).  With the configs' weak scaling (``n_clips = world * B_local``) this is ``[r*B_local, (r+1)*B_local)``."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n_clips < 0:
        raise ValueError("n_clips must be non-negative")
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's clips of a batch-major tensor (a view, no copy)."""
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def allreduce_grads_(grads: Sequence[torch.Tensor], world: int | None = None, mean: bool = False, group=None) -> None:
    """In-place SUM (or mean) over ranks of a list of gradient tensors through ONE flat buffer (one collective launch
    instead of one per tensor: the payload is small -- 3.01 M floats for the generator -- so latency, not bandwidth,
    is what matters on NVSwitch)."""
    grads = [g for g in grads if g is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if mean:
        flat.div_(world or dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def max_over_ranks(value: float, device: torch.device | str = "cpu", group=None) -> float:
    """Timing rule of the bench: a multi-GPU number is the MAX over ranks of the device-measured time."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def whole_job_throughput(units_per_rank: Iterable[float], seconds: float) -> float:
    """``value`` of the bench line: the units ALL ranks processed divided by the max-over-ranks time."""
    return float(sum(units_per_rank)) / seconds
