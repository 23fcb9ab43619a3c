"""Attention core alone (hdrvae_attention: q, k, v 16-bit [B, T, 512]) — device time and effective TFLOP/s
(4 * T^2 * 512 per image).  The transposes / interleaves of the test entry are excluded by timing a second run's
profile scope... simpler: time the whole entry for large T where they are < 1 %.
  HDRVAE_ATTN_FUSED=0|1 python tools/attn_bench.py"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict  # noqa: E402

dev = torch.device("cuda:0")
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
cases = [(4, 4096), (4, 16384), (1, 65536), (1, 262144)]
if len(sys.argv) > 1:
    cases = [(int(a.split("x")[0]), int(a.split("x")[1])) for a in sys.argv[1:]]
for B, T in cases:
    g = torch.Generator(device=dev).manual_seed(T)
    q = (torch.randn(B, T, 512, generator=g, device=dev) * 2).half()
    k = torch.randn(B, T, 512, generator=g, device=dev).half()
    v = torch.randn(B, T, 512, generator=g, device=dev).half()
    o = eng.attention(q, k, v)
    torch.cuda.synchronize()
    reps = 3 if T >= 65536 else 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        o = eng.attention(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 4.0 * B * T * T * 512 / (ms / 1e3) / 1e12
    print(f"fused={os.environ.get('HDRVAE_ATTN_FUSED', '1')} cg={os.environ.get('HDRVAE_ATTN_CG', '-')} B={B} T={T}: {ms:.3f} ms  {tf:.1f} TFLOP/s effective  finite={bool(torch.isfinite(o).all())}", flush=True)
