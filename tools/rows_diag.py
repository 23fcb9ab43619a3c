"""Row-tiling diagnostics: per-row error profile of the emulated tiled decode against the single-GPU decode."""
import sys

import torch

sys.path.insert(0, ".")
from oracle.flux_decoder import build_decoder, make_latent  # noqa: E402
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.sharding import decode_rows_emulated  # noqa: E402

dev = "cuda:0"
eng = HdrVaeEngine(build_decoder(0).state_dict(), dev)
h, w = 16, 16
z = make_latent(1, h, w, seed=43).to(dev)
whole, _ = eng.decode(z, "conservative")
whole2, _ = eng.decode(z + 0.0, "conservative")
print("whole vs whole (determinism):", float((whole - whole2).abs().max()))
for world in (1, 2, 4):
    tiled, _ = decode_rows_emulated(eng, z, world, "conservative")
    d = (tiled - whole).double()
    rel = float(d.norm() / whole.double().norm())
    err = (d ** 2).mean(dim=(0, 2, 3)).sqrt()
    slab = 8 * h // world
    print(f"world={world}: rel {rel:.3e}; per-row rms (x1e4), seams at multiples of {slab}:")
    print("   ", " ".join(f"{1e4 * float(e):.1f}" for e in err))
