"""Config C5 end to end on one GPU: 1x16x128x128 latent -> HDR decode 1024^2 -> 4x HDR upscale (random-init ESRGAN,
nb 23) -> half packing in EXR scanline order.  Device-timed per stage and in total."""
import sys

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200.engine import HdrVaeEngine, pack_half  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict, random_upscaler_state_dict, synthetic_latent  # noqa: E402
from vae_decode_hdr_b200.upscaler import HdrUpscalerEngine  # noqa: E402

dev = torch.device("cuda:0")
dec = HdrVaeEngine(random_decoder_state_dict(0), dev)
up = HdrUpscalerEngine(random_upscaler_state_dict(0, 23), dev)
z = synthetic_latent(1, 128, 128).to(dev)


def run():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    img, _ = dec.decode(z, "moderate", want_stats=False)
    ev[1].record()
    big = up.upscale(img, "atanh")
    ev[2].record()
    half = pack_half(big, exr_scanline_order=True)
    ev[3].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)], half


for _ in range(2):
    run()
ts = [run()[0] for _ in range(3)]
t = [sum(x[i] for x in ts) / len(ts) for i in range(3)]
_, half = run()
print(f"C5: decode 1024^2 {t[0]:.2f} ms + upscale 4x {t[1]:.2f} ms + half pack {t[2]:.2f} ms = {sum(t):.2f} ms "
      f"({16.78 / (sum(t) / 1e3):.1f} output MP/s); packed {tuple(half.shape)} {half.dtype}, finite {bool(torch.isfinite(half.float()).all())}")
