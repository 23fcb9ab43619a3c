"""Row-tiling diagnostics: at every HALO exchange compare the exchanged tensor of a world=R emulation with world=1."""
import sys

import torch

sys.path.insert(0, ".")
from oracle.flux_decoder import build_decoder, make_latent  # noqa: E402
from vae_decode_hdr_b200 import _native as N  # noqa: E402
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200 import sharding as S  # noqa: E402

dev = "cuda:0"
eng = HdrVaeEngine(build_decoder(0).state_dict(), dev)
h, w = int(sys.argv[1]), int(sys.argv[2])
WORLD = int(sys.argv[3])
MODE = sys.argv[4] if len(sys.argv) > 4 else "conservative"
z = make_latent(1, h, w, seed=int(sys.argv[5]) if len(sys.argv) > 5 else 43).to(dev)


def run(world):
    states, wss, outs, keep = [], [], [], []
    for r in range(world):
        st, ws, out, zz = eng.rows_begin(z, r, world, MODE, 1.0)
        states.append(st); wss.append(ws); outs.append(out); keep.append(zz)
    trace = []
    while True:
        exs = [eng.rows_run(st) for st in states]
        ex = exs[0]
        if ex.kind == N.EX_END:
            break
        if ex.kind & N.EX_HALO:
            i = 0
            n = ex.halo_row_bytes[i]
            rows = (ex.halo_last_row_off[i] - ex.halo_first_row_off[i]) // n + 1
            t = torch.cat([ws[ex.halo_first_row_off[i]:ex.halo_first_row_off[i] + rows * n].view(torch.float32).clone() for ws in wss])
            sums = wss[0][ex.allreduce_off:ex.allreduce_off + 8 * ex.allreduce_count].view(torch.float64).clone() if ex.kind & N.EX_ALLREDUCE_F64 else None
            trace.append((rows * world, n, t, torch.stack([ws[ex.allreduce_off:ex.allreduce_off + 8 * ex.allreduce_count].view(torch.float64) for ws in wss]).sum(0).clone()))
        # perform the exchange exactly as sharding.decode_rows_emulated does
        if ex.kind & N.EX_HALO:
            for i in range(ex.n_halo):
                n = ex.halo_row_bytes[i]
                for r in range(world):
                    if r > 0:
                        wss[r - 1][ex.halo_bottom_off[i]:ex.halo_bottom_off[i] + n].copy_(wss[r][ex.halo_first_row_off[i]:ex.halo_first_row_off[i] + n])
                    if r < world - 1:
                        wss[r + 1][ex.halo_top_off[i]:ex.halo_top_off[i] + n].copy_(wss[r][ex.halo_last_row_off[i]:ex.halo_last_row_off[i] + n])
        if ex.kind & N.EX_ALLREDUCE_F64:
            views = [ws[ex.allreduce_off:ex.allreduce_off + 8 * ex.allreduce_count].view(torch.float64) for ws in wss]
            total = torch.stack(views).sum(0)
            for v in views:
                v.copy_(total)
        if ex.kind & N.EX_ALLGATHER:
            for i in range(ex.n_gather):
                n = ex.gather_bytes_per_rank[i]
                parts = [wss[r][ex.gather_off[i] + r * n:ex.gather_off[i] + (r + 1) * n].clone() for r in range(world)]
                for ws in wss:
                    for r in range(world):
                        ws[ex.gather_off[i] + r * n:ex.gather_off[i] + (r + 1) * n].copy_(parts[r])
        if ex.kind & N.EX_RAW_STATS:
            blocks = [S._raw_views(ws, ex.raw_stats_off) for ws in wss]
            vmin, vmax, vsum = S.merge_raw_stats(blocks)
            for b in blocks:
                b[0].copy_(vmin); b[1].copy_(vmax); b[2].copy_(vsum)
    for st in states:
        eng.rows_end(st, False)
    return trace, torch.cat(outs, dim=1)


t1, o1 = run(1)
t2, o2 = run(WORLD)
print("steps", len(t1), len(t2))
for k, (a, b) in enumerate(zip(t1, t2)):
    rel = float((a[2] - b[2]).double().norm() / a[2].double().norm())
    srel = float((a[3] - b[3]).abs().max() / a[3].abs().max())
    print(f"halo step {k:2d}: rows {a[0]:4d} row_bytes {a[1]:8d}  tensor rel diff {rel:.3e}   GN sums rel diff {srel:.3e}")
err = ((o1 - o2).double() ** 2).mean(dim=(0, 2, 3)).sqrt()
print("per-row rms x1e4:", " ".join(f"{float(e)*1e4:.1f}" for e in err))
print("final image rel", float((o1 - o2).double().norm() / o1.double().norm()))
