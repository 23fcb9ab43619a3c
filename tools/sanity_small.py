"""Small end-to-end run of every product path (for compute-sanitizer): decode (all modes), row-tiled emulation, upscale."""
import sys

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200.engine import HdrVaeEngine, pack_half, quantiles  # noqa: E402
from vae_decode_hdr_b200.sharding import decode_rows_emulated  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict, random_upscaler_state_dict, synthetic_latent  # noqa: E402
from vae_decode_hdr_b200.upscaler import HdrUpscalerEngine  # noqa: E402

dev = torch.device("cuda:0")
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
for shape in [(1, 16, 16), (2, 5, 9)]:
    z = synthetic_latent(*shape).to(dev)
    for mode in ("conservative", "exposure", "adaptive_recovery", "mathematical_recovery"):
        out, st = eng.decode(z, mode)
    print("decode", shape, tuple(out.shape), st["out_max"])
z = synthetic_latent(1, 32, 8).to(dev)
out, st = decode_rows_emulated(eng, z, 2, "moderate")
print("rows", tuple(out.shape), st["out_max"])
print("quantiles", quantiles(out, (0.5, 0.99)))
print("half", tuple(pack_half(out, True).shape))
up = HdrUpscalerEngine(random_upscaler_state_dict(0, 2), dev)
img = torch.rand(1, 530, 24, 3, device=dev) * 2
big = up.upscale(img, "atanh", True, True, "bilinear")
print("upscale", tuple(big.shape), float(big.abs().max()))
torch.cuda.synchronize()
print("done")
