"""A/B of epilogue variants on one box (env HDRVAE_GEMM_DBG bit 2 = un-pipelined TMEM loads) for decoder-shaped convs
with GroupNorm statistics, with and without the fp32 residual."""
import os
import sys

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict  # noqa: E402

dev = "cuda:0"
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
for name, B, H, cin, cout in [("128->128@1024", 4, 1024, 128, 128), ("256->256@512", 4, 512, 256, 256), ("512->512@256", 4, 256, 512, 512)]:
    x = torch.randn(B, H, H, cin, device=dev).half()
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.02
    b = torch.zeros(cout, device=dev)
    r = torch.randn(B, H, H, cout, device=dev)
    for res in (None, r):
        line = f"{name:14s} {'res+stats' if res is not None else 'stats    '}:"
        for rep in range(2):
            for dbg in (0, 4):
                os.environ["HDRVAE_GEMM_DBG"] = str(dbg)
                eng.lib.hdrvae_profile_begin()
                for _ in range(5):
                    eng.conv2d(x, w, b, 3, False, res, out_dtype=torch.float32, want_stats=True)
                path = f"/tmp/epi_{dbg}.tsv"
                eng.lib.hdrvae_profile_end(path.encode())
                ms = sorted(float(l.split("\t")[1].split()[0]) for l in open(path) if l.startswith("conv"))
                line += f"  {'pipe' if dbg == 0 else 'flat'} {ms[len(ms) // 2]:.4f}"
        print(line)
os.environ["HDRVAE_GEMM_DBG"] = "0"
