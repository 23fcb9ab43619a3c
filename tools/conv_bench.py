"""Micro-benchmark of the tcgen05 conv kernel on the decoder's layer shapes (CUDA events, L2 flushed
between launches by cycling through distinct input buffers larger than L2).  Usage:
  python tools/conv_bench.py [B] [latent]     (default 4 128: BASELINE config C2)"""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle.flux_decoder import build_decoder  # noqa: E402
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = "cuda:0"
eng = HdrVaeEngine(build_decoder(0).state_dict(), dev)
shapes = [  # (name, H, Cin, Cout, ks, upsample)
    ("res512@1x", L, 512, 512, 3, False), ("res512@2x", 2 * L, 512, 512, 3, False), ("up512@1x", L, 512, 512, 3, True),
    ("up512@2x", 2 * L, 512, 512, 3, True), ("res512->256@4x", 4 * L, 512, 256, 3, False),
    ("res256@4x", 4 * L, 256, 256, 3, False), ("up256@4x", 4 * L, 256, 256, 3, True),
    ("res256->128@8x", 8 * L, 256, 128, 3, False), ("res128@8x", 8 * L, 128, 128, 3, False),
    ("nin256->128@8x", 8 * L, 256, 128, 1, False), ("qkv512@1x", L, 512, 512, 1, False),
]
print(f"B={B} latent={L}")
for name, H, cin, cout, ks, up in shapes:
    x = torch.randn(B, H, H, cin, device=dev).bfloat16()
    w = torch.randn(cout, cin, ks, ks, device=dev) * 0.02
    b = torch.zeros(cout, device=dev)
    OH = 2 * H if up else H
    flops = 2.0 * B * OH * OH * cout * cin * ks * ks          # algorithmic (3x3 on the upsampled grid)
    for _ in range(2):
        y = eng.conv2d(x, w, b, ks, up)
    torch.cuda.synchronize()
    # hdrvae_conv2d repacks weights per call (test entry); time only the conv kernels via events around 3 calls
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 3
    t0 = time.time()
    ev[0].record()
    for _ in range(n):
        y = eng.conv2d(x, w, b, ks, up)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    print(f"{name:18s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s (incl. weight repack + sync of the test entry)")
    del x, y
