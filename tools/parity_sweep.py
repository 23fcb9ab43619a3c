"""Whole-decode parity sweep: rel-L2 of the 16-bit CUDA decode vs the fp32 oracle over sizes, modes and seeds."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import hdr_oracle as ho  # noqa: E402
from oracle.flux_decoder import build_decoder, make_latent  # noqa: E402
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402

dev = "cuda:0"
dec = build_decoder(0)
eng = HdrVaeEngine(dec.state_dict(), dev)
modes = ["conservative", "moderate", "exposure", "adaptive_recovery", "mathematical_recovery"]
print("size seed " + " ".join(f"{m[:12]:>12s}" for m in modes) + "   features")
for (h, w) in [(8, 8), (8, 12), (16, 16), (32, 32)]:
    for seed in (41, 45, 7):
        z = make_latent(1, h, w, seed=seed)
        row = []
        for m in modes:
            ref, _, _ = ho.simple_hdr_decode(dec, z, m, 1.0)
            out, _ = eng.decode(z.to(dev), m)
            row.append(float((out.cpu() - ref).double().norm() / ref.double().norm()))
        f = eng.decode_features(z.to(dev)).float().cpu()
        fr = dec.features(z).permute(0, 2, 3, 1)
        frel = float((f - fr).double().norm() / fr.double().norm())
        print(f"{h}x{w} {seed:3d} " + " ".join(f"{r:12.3e}" for r in row) + f"   {frel:.3e}")
