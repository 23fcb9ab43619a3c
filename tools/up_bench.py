"""Config C5 timing: 4x HDR upscale of one image with a random-init ESRGAN (RRDBNet nf 64, nb 23, gc 32).
python tools/up_bench.py [H] [W] [nb] [profile.tsv]"""
import sys

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200 import _native  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_upscaler_state_dict  # noqa: E402
from vae_decode_hdr_b200.upscaler import HdrUpscalerEngine  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 23
out = sys.argv[4] if len(sys.argv) > 4 else None
dev = torch.device("cuda:0")
eng = HdrUpscalerEngine(random_upscaler_state_dict(0, nb), dev)
g = torch.Generator().manual_seed(1)
img = (torch.rand(1, H, W, 3, generator=g) * 2.5).to(dev)
lib = _native.load_library()
res = eng.upscale(img)
torch.cuda.synchronize()
print("out", tuple(res.shape), "range", float(res.min()), float(res.max()), "finite", bool(torch.isfinite(res).all()))
lib.hdrvae_profile_begin()
eng.upscale(img)
lib.hdrvae_profile_end(out.encode() if out else None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.upscale(img)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
# model FLOPs per input pixel of a tile (2 * MACs): conv_first + nb * 3 RDBs + conv_body + up1 (4 px) + up2, hr, last (16 px)
rdb = 18 * (64 * 32 + 96 * 32 + 128 * 32 + 160 * 32 + 192 * 64)
per_px = 18 * 3 * 64 + nb * 3 * rdb + 18 * 64 * 64 * (1 + 4 + 16 + 16) + 16 * 18 * 64 * 3
pos = lambda n: [0] if n <= 512 else list(range(0, n - 64, 448))  # noqa: E731
tile_px = sum(min(512, H - y) * min(512, W - x) for y in pos(H) for x in pos(W))
flops = 2 * tile_px * per_px
print(f"upscale {H}x{W} -> {4*H}x{4*W} nb={nb}: {ms:.2f} ms, {flops/1e12:.1f} TFLOP algorithmic (2 passes, {tile_px/H/W:.2f}x tile overlap) "
      f"= {flops/ms/1e9:.0f} TFLOP/s, {16*H*W/1e6/(ms/1e3):.1f} output MP/s")
