"""Config C3 per-GPU shapes on one GPU: B x16x128x128 latents -> B x 1024^2, exposure mode (B = 32 on 1 GPU, 16 / 8 / 4
per GPU when batch-sharded over 2 / 4 / 8).  python tools/c3_batch.py [B ...]"""
import sys

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict, synthetic_latent  # noqa: E402

dev = torch.device("cuda:0")
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
for B in [int(a) for a in sys.argv[1:]] or [32, 16, 8, 4]:
    z = synthetic_latent(B, 128, 128).to(dev)
    ws = eng.workspace_bytes(B, 128, 128) / 2**30
    for _ in range(2):
        out, st = eng.decode(z, "exposure")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out, _ = eng.decode(z, "exposure", want_stats=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"C3 B={B}: workspace {ws:.1f} GiB, {ms:.1f} ms/step, {B * 1.048576 / (ms / 1e3):.1f} MP/s, out max {st['out_max']:.2f}, "
          f"hdr_pixels {st['hdr_pixels']}, finite {bool(torch.isfinite(out).all())}")
    del out, z
    eng._workspace = None
    torch.cuda.empty_cache()
