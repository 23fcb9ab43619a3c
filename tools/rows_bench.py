"""Row-tiled decode of ONE large image across the ranks of a torchrun job (config C4 path): device-timed (max over
ranks), checked against the single-GPU decode of the same latent on rank 0 when --check is given.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/rows_bench.py 512 512 --steps 5"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.sharding import RowsDirect, decode_rows_sharded  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict, synthetic_latent  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("h", type=int)
ap.add_argument("w", type=int)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--mode", default="moderate")
ap.add_argument("--check", action="store_true")
ap.add_argument("--transport", default="nccl", choices=["nccl", "p2p"])
a = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
z = synthetic_latent(1, a.h, a.w, seed=3).to(dev)

p2p = RowsDirect(eng, a.h, a.w) if a.transport == "p2p" else None


def run(want_stats):
    if p2p is not None:
        return p2p.decode(z, a.mode, 1.0, want_stats=want_stats)
    return decode_rows_sharded(eng, z, a.mode, 1.0, want_stats=want_stats)


for _ in range(a.warmup):
    out, st = run(True)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out, st = run(False)
e1.record()
torch.cuda.synchronize(); dist.barrier()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
res = {"workload": f"1x16x{a.h}x{a.w} latent -> {8*a.h}x{8*a.w}, {a.mode}, row-tiled over {world} GPUs", "n_gpus": world,
       "transport": a.transport, "ms_per_image": float(ms), "megapixels_per_s": 64.0 * a.h * a.w / 1e6 / (float(ms) / 1e3),
       "peak_mem_gib_per_gpu": torch.cuda.max_memory_allocated() / 2**30}
if a.check:
    rows = 8 * a.h // world
    parts = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
    dist.gather(out.contiguous(), parts, dst=0)
    if rank == 0:
        del out
        tiled = torch.cat(parts, dim=1)
        del parts
        torch.cuda.empty_cache()
        whole, _ = eng.decode(z, a.mode, 1.0)
        d = 0.0
        n = 0.0
        for r0 in range(0, whole.shape[1], 256):       # chunked: fp64 copies of a 4096^2 image are large
            wch, tch = whole[:, r0:r0 + 256].double(), tiled[:, r0:r0 + 256].double()
            d += float(((wch - tch) ** 2).sum()); n += float((wch ** 2).sum())
        res["rel_l2_vs_single_gpu"] = (d / n) ** 0.5
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
