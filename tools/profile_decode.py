"""Per-op device timing of one decode (library profiler scopes).  python tools/profile_decode.py [B] [latent] [out.tsv]"""
import sys

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200 import _native  # noqa: E402
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict, synthetic_latent  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
out = sys.argv[3] if len(sys.argv) > 3 else None
dev = torch.device("cuda:0")
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
z = synthetic_latent(B, L, L).to(dev)
lib = _native.load_library()
for _ in range(2):
    eng.decode(z, "moderate", want_stats=False)
torch.cuda.synchronize()
lib.hdrvae_profile_begin()
img, st = eng.decode(z, "moderate")
lib.hdrvae_profile_end(out.encode() if out else None)
print({k: st[k] for k in ("pre_min", "pre_max", "pre_mean", "post_min", "post_max", "out_max", "hdr_pixels", "norm_function", "accepted")})
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.decode(z, "moderate", want_stats=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"decode B={B} latent={L}: {ms:.2f} ms/step, {B * (8 * L) ** 2 / 1e6 / (ms / 1e3):.1f} MP/s")
