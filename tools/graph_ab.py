"""Device-timed C2 steps (CUDA events around N decodes) — run with and without HDRVAE_NO_GRAPH=1 to see what the graph buys."""
import sys

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict, synthetic_latent  # noqa: E402

dev = torch.device("cuda:0")
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
z = synthetic_latent(4, 128, 128).to(dev)
for _ in range(6):
    eng.decode(z, "moderate", want_stats=False)
torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.decode(z, "moderate", want_stats=False)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 10)
n0 = eng.lib.hdrvae_launch_count()
eng.decode(z, "moderate", want_stats=False)
torch.cuda.synchronize()
print("ms/step", [f"{r:.2f}" for r in res], "launches per step", eng.lib.hdrvae_launch_count() - n0)
