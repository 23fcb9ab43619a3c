"""Determinism soak: the same input must give bit-identical output on every repeat, across shapes, modes, the CUDA-graph
path, the row-tiled emulation and the upscaler (a barrier-protocol race in the tcgen05 kernels would show up here)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.sharding import decode_rows_emulated  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict, random_upscaler_state_dict, synthetic_latent  # noqa: E402
from vae_decode_hdr_b200.upscaler import HdrUpscalerEngine  # noqa: E402

dev = torch.device("cuda:0")
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
t0 = time.time()
bad = 0
shapes = [(1, 16, 16), (2, 5, 9), (4, 32, 32), (1, 64, 48), (4, 128, 128), (1, 24, 40)]
modes = ["moderate", "exposure", "adaptive_recovery", "aggressive"]
first = {}
for rep in range(12):
    for i, shp in enumerate(shapes):
        z = synthetic_latent(*shp, seed=100 + i).to(dev)
        mode = modes[(i + rep) % len(modes)]
        out, _ = eng.decode(z, mode, want_stats=(rep % 3 == 0))
        key = (shp, mode)
        if key not in first:
            first[key] = out.clone()
        elif not torch.equal(out, first[key]):
            bad += 1
            print("MISMATCH decode", key, "rep", rep, float((out - first[key]).abs().max()))
print(f"decode: {12 * len(shapes)} runs, {len(first)} distinct (shape, mode), mismatches {bad}")
z = synthetic_latent(1, 32, 24, seed=7).to(dev)
ref = None
for rep in range(6):
    out, _ = decode_rows_emulated(eng, z, 2 + 2 * (rep % 2), "moderate")       # 2 and 4 virtual ranks alternate
    k = rep % 2
    if rep < 2:
        ref = ref or {}
        ref[k] = out.clone()
    elif not torch.equal(out, ref[k]):
        bad += 1
        print("MISMATCH rows", rep)
print("rows emulation: 6 runs, mismatches so far", bad)
up = HdrUpscalerEngine(random_upscaler_state_dict(0, 3), dev)
img = (torch.rand(1, 530, 72, 3, device=dev) * 2.5)
u0 = up.upscale(img).clone()
for rep in range(8):
    if not torch.equal(up.upscale(img), u0):
        bad += 1
        print("MISMATCH upscale", rep)
torch.cuda.synchronize()
print(f"soak done in {time.time() - t0:.1f} s, total mismatches {bad}")
sys.exit(1 if bad else 0)
