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


# ------------------------------------------------------------------------------------------------ row tiling
# One image, latent rows split evenly over the ranks (BASELINE config C4).  libhdrvae runs the same step program
# on every rank's slab and tells the host what to exchange between steps (include/hdrvae.h, hdrvae_exchange):
# 1-row conv halos to the neighbours, GroupNorm sums all-reduced, attention K/V all-gathered, HDR statistics
# all-reduced.  Here the exchanges are NCCL collectives / point-to-point (torch.distributed) on views of the
# workspace; `decode_rows_emulated` performs the same exchanges with plain copies between R workspaces of ONE
# process, which is how the path is tested on a single GPU.
from . import _native as _N  # noqa: E402


def _raw_views(ws: torch.Tensor, off: int):
    blk = ws[off:off + 96]
    return blk[0:16].view(torch.float32), blk[16:32].view(torch.float32), blk[32:96].view(torch.float64)


def _exchange_nccl(ex, ws: torch.Tensor, rank: int, world: int, group=None) -> None:
    if ex.kind & _N.EX_HALO:
        ops = []
        for i in range(ex.n_halo):
            n = ex.halo_row_bytes[i]
            first, last = ws[ex.halo_first_row_off[i]:ex.halo_first_row_off[i] + n], ws[ex.halo_last_row_off[i]:ex.halo_last_row_off[i] + n]
            top, bot = ws[ex.halo_top_off[i]:ex.halo_top_off[i] + n], ws[ex.halo_bottom_off[i]:ex.halo_bottom_off[i] + n]
            if rank > 0:
                ops += [dist.P2POp(dist.isend, first, rank - 1, group), dist.P2POp(dist.irecv, top, rank - 1, group)]
            if rank < world - 1:
                ops += [dist.P2POp(dist.isend, last, rank + 1, group), dist.P2POp(dist.irecv, bot, rank + 1, group)]
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
    if ex.kind & _N.EX_ALLREDUCE_F64:
        buf = ws[ex.allreduce_off:ex.allreduce_off + 8 * ex.allreduce_count].view(torch.float64)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    if ex.kind & _N.EX_ALLGATHER:
        for i in range(ex.n_gather):
            n = ex.gather_bytes_per_rank[i]
            full = ws[ex.gather_off[i]:ex.gather_off[i] + n * world]
            dist.all_gather_into_tensor(full, full[rank * n:(rank + 1) * n].clone(), group=group)
    if ex.kind & _N.EX_RAW_STATS:
        allreduce_raw_stats(*_raw_views(ws, ex.raw_stats_off), group=group)


def decode_rows_sharded(engine, latent_full: torch.Tensor, hdr_mode: str, ev_multiplier: float = 1.0, group=None,
                        want_stats: bool = True):
    """Row-tiled decode of ONE image across the ranks of `group`: returns this rank's rows of the image
    ([1, 8h/world, 8w, 3]) and the statistics (pre/post/conv/pre3 global, out_* of the local slab)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    state, ws, out, _keep = engine.rows_begin(latent_full, rank, world, hdr_mode, ev_multiplier)
    while True:
        ex = engine.rows_run(state)
        if ex.kind == _N.EX_END:
            break
        _exchange_nccl(ex, ws, rank, world, group)
    return out, engine.rows_end(state, want_stats)


def decode_rows_emulated(engine, latent_full: torch.Tensor, world: int, hdr_mode: str, ev_multiplier: float = 1.0):
    """The same row-tiled program with `world` virtual ranks on ONE GPU (one workspace each, run in lock step);
    exchanges are device copies.  Returns the full image [1, 8h, 8w, 3] and rank 0's statistics."""
    states, wss, outs, keep = [], [], [], []
    for r in range(world):
        st, ws, out, z = engine.rows_begin(latent_full, r, world, hdr_mode, ev_multiplier)
        states.append(st); wss.append(ws); outs.append(out); keep.append(z)
    while True:
        exs = [engine.rows_run(st) for st in states]
        ex = exs[0]
        if ex.kind == _N.EX_END:
            break
        if ex.kind & _N.EX_HALO:
            for i in range(ex.n_halo):
                n = ex.halo_row_bytes[i]
                for r in range(world):
                    if r > 0:       # my first interior row -> upper neighbour's bottom halo
                        wss[r - 1][ex.halo_bottom_off[i]:ex.halo_bottom_off[i] + n].copy_(wss[r][ex.halo_first_row_off[i]:ex.halo_first_row_off[i] + n])
                    if r < world - 1:   # my last interior row -> lower neighbour's top halo
                        wss[r + 1][ex.halo_top_off[i]:ex.halo_top_off[i] + n].copy_(wss[r][ex.halo_last_row_off[i]:ex.halo_last_row_off[i] + n])
        if ex.kind & _N.EX_ALLREDUCE_F64:
            views = [ws[ex.allreduce_off:ex.allreduce_off + 8 * ex.allreduce_count].view(torch.float64) for ws in wss]
            total = torch.stack(views).sum(0)
            for v in views:
                v.copy_(total)
        if ex.kind & _N.EX_ALLGATHER:
            for i in range(ex.n_gather):
                n = ex.gather_bytes_per_rank[i]
                parts = [wss[r][ex.gather_off[i] + r * n:ex.gather_off[i] + (r + 1) * n].clone() for r in range(world)]
                for ws in wss:
                    for r in range(world):
                        ws[ex.gather_off[i] + r * n:ex.gather_off[i] + (r + 1) * n].copy_(parts[r])
        if ex.kind & _N.EX_RAW_STATS:
            blocks = [_raw_views(ws, ex.raw_stats_off) for ws in wss]
            vmin, vmax, vsum = merge_raw_stats(blocks)
            for b in blocks:
                b[0].copy_(vmin); b[1].copy_(vmax); b[2].copy_(vsum)
    stats = [engine.rows_end(st, True) for st in states]
    return torch.cat(outs, dim=1), stats[0]


class RowsP2P:
    """Row-tiled decode with the conv halos pushed straight into the neighbours' workspaces over NVLink peer-to-peer
    (CUDA IPC mappings of the neighbour ranks' workspace; the library issues one stream-ordered peer copy per halo row
    right after the conv that produced it: hdrvae_rows_set_peers) instead of NCCL send/recv.  The GroupNorm all-reduce that accompanies every halo exchange doubles as the synchronisation: a
    rank's all-reduce kernel is stream-ordered after its pushes, so when the all-reduce completes on a rank every
    neighbour's push into that rank has landed; two writes of the same halo row are always separated by at least one
    all-reduce, so a push can never overtake the neighbour's last read of the previous contents.  One NCCL call per
    exchange point instead of three.  K/V all-gather and the HDR statistics stay on NCCL.

    The mappings use torch's CUDA-IPC plumbing (`UntypedStorage._share_cuda_` / `_new_shared_cuda`, the private calls
    behind torch.multiprocessing's tensor sharing) — the one place this package leans on a private torch API; the C ABI
    itself only sees raw pointers (any cudaIpcOpenMemHandle mapping will do).
    The workspace is persistent (the IPC handles are exchanged once, in the constructor); `decode` may be called any
    number of times for latents of the shape given at construction."""

    def __init__(self, engine, h: int, w: int, group=None):
        self.engine, self.group, self.h, self.w = engine, group, h, w
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        need = engine.rows_workspace_bytes(h, w, self.world)
        with torch.cuda.device(engine.device):
            self.ws = torch.empty(need, dtype=torch.uint8, device=engine.device)
            torch.cuda.synchronize(engine.device)
        info = (self.ws.untyped_storage()._share_cuda_(), self.ws.storage_offset(), need)
        infos = [None] * self.world
        dist.all_gather_object(infos, info, group=group)
        self.peers = {}
        for r in (self.rank - 1, self.rank + 1):
            if 0 <= r < self.world:
                share, offset, size = infos[r]
                storage = torch.UntypedStorage._new_shared_cuda(*share)
                t = torch.empty(0, dtype=torch.uint8, device=torch.device("cuda", share[0])).set_(storage)
                self.peers[r] = t[offset:offset + size]
        dist.barrier(group=group)       # nobody frees or reuses its workspace before every mapping exists

    def _exchange(self, ex) -> None:
        ws, rank, world = self.ws, self.rank, self.world
        if ex.kind & _N.EX_HALO:
            raise RuntimeError("the library did not push the halo rows (hdrvae_rows_set_peers not in effect)")
        if ex.kind & _N.EX_ALLREDUCE_F64:
            buf = ws[ex.allreduce_off:ex.allreduce_off + 8 * ex.allreduce_count].view(torch.float64)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        elif ex.kind & _N.EX_HALO_PUSHED:
            dist.barrier(group=self.group)      # no all-reduce at this point: explicit synchronisation
        if ex.kind & _N.EX_ALLGATHER:
            for i in range(ex.n_gather):
                n = ex.gather_bytes_per_rank[i]
                full = ws[ex.gather_off[i]:ex.gather_off[i] + n * world]
                dist.all_gather_into_tensor(full, full[rank * n:(rank + 1) * n].clone(), group=self.group)
        if ex.kind & _N.EX_RAW_STATS:
            allreduce_raw_stats(*_raw_views(ws, ex.raw_stats_off), group=self.group)

    def decode(self, latent_full: torch.Tensor, hdr_mode: str, ev_multiplier: float = 1.0, want_stats: bool = True):
        if tuple(latent_full.shape[-2:]) != (self.h, self.w):
            raise ValueError(f"this RowsP2P was built for {self.h}x{self.w} latents, got {tuple(latent_full.shape)}")
        eng = self.engine
        state, _ws, out, _keep = eng.rows_begin(latent_full, self.rank, self.world, hdr_mode, ev_multiplier, workspace=self.ws)
        up = self.peers[self.rank - 1].data_ptr() if self.rank > 0 else None
        down = self.peers[self.rank + 1].data_ptr() if self.rank < self.world - 1 else None
        _N.check(eng.lib.hdrvae_rows_set_peers(state, up, down), "hdrvae_rows_set_peers")
        while True:
            ex = eng.rows_run(state)
            if ex.kind == _N.EX_END:
                break
            self._exchange(ex)
        return out, eng.rows_end(state, want_stats)
