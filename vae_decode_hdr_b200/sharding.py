"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Batch sharding (BASELINE config C3): images are independent through the decoder
(GroupNorm is per sample), but every global statistic of the HDR epilogue is taken over the WHOLE
batch (hdr_vae_decode.py:862-865,1098,1116).  So the path has exactly one exchange
step: between epilogue phase A and phase B the raw statistics block (4 mins, 4 maxes, 8 sums) is
all-reduced, after which every rank finishes its own images with identical scalars — results do
not depend on the number of GPUs."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [start, end) ranges; the first n_items % world ranks get one extra."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    base, extra = divmod(n_items, world_size)
    out, s = [], 0
    for r in range(world_size):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def allreduce_raw_stats(vmin: torch.Tensor, vmax: torch.Tensor, vsum: torch.Tensor, group=None) -> None:
    """In-place cross-rank reduction of an hdrvae_raw_stats block (include/hdrvae.h): MIN / MAX / SUM."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(vmin, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(vsum, op=dist.ReduceOp.SUM, group=group)


def merge_raw_stats(blocks: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]):
    """Single-process equivalent of :func:`allreduce_raw_stats` (tests / emulation)."""
    vmin = torch.stack([b[0] for b in blocks]).amin(0)
    vmax = torch.stack([b[1] for b in blocks]).amax(0)
    vsum = torch.stack([b[2] for b in blocks]).sum(0)
    return vmin, vmax, vsum


def decode_batch_sharded(engine, latent_local: torch.Tensor, hdr_mode: str, ev_multiplier: float = 1.0, group=None,
                         want_stats: bool = True):
    """Decode this rank's slice of the batch with batch-global HDR statistics.

    engine: vae_decode_hdr_b200.engine.HdrVaeEngine on this rank's GPU.  Returns (image_local, stats)
    where the pre/post/conv/pre3 statistics and every derived scalar are those of the whole batch;
    out_min/out_max/hdr_pixels describe the local slice."""
    vmin, vmax, vsum = engine.decode_begin(latent_local)
    allreduce_raw_stats(vmin, vmax, vsum, group)
    return engine.decode_finish(hdr_mode, ev_multiplier, want_stats)
