"""Runs one decoder-shaped conv a few times (for ncu captures).  python tools/one_conv.py H Cin Cout ks [res] [up] [B]"""
import sys

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict  # noqa: E402

H, cin, cout, ks = (int(a) for a in sys.argv[1:5])
res = len(sys.argv) > 5 and sys.argv[5] == "1"
up = len(sys.argv) > 6 and sys.argv[6] == "1"
B = int(sys.argv[7]) if len(sys.argv) > 7 else 1
dev = "cuda:0"
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
x = torch.randn(B, H, H, cin, device=dev).half()
w = torch.randn(cout, cin, ks, ks, device=dev) * 0.02
b = torch.zeros(cout, device=dev)
OH = 2 * H if up else H
r = torch.randn(B, OH, OH, cout, device=dev) if res else None
for i in range(4):
    y, part = eng.conv2d(x, w, b, ks, up, r, out_dtype=torch.float32, want_stats=True)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.float().abs().mean()))
