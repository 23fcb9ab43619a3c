"""One large decode (BASELINE config C4: 1x16x512x512 -> 4096x4096, 'aggressive') on one GPU: time, and parity
against the fp32 PyTorch oracle on the same GPU when --check is given.  python tools/big_decode.py [latent] [--check]"""
import sys
import time

import torch

sys.path.insert(0, ".")
L = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 512
check = "--check" in sys.argv
dev = torch.device("cuda:0")
from oracle.flux_decoder import build_decoder, make_latent  # noqa: E402  (checker only)
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402

dec = build_decoder(0)
eng = HdrVaeEngine(dec.state_dict(), dev)
z = make_latent(1, L, L, seed=1234).to(dev)
print(f"latent {L}x{L}: workspace {eng.workspace_bytes(1, L, L) / 2**30:.1f} GiB")
out, st = eng.decode(z, "aggressive")
torch.cuda.synchronize()
ts = []
for _ in range(2):
    t0 = time.perf_counter()
    out, st = eng.decode(z, "aggressive")
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
mp = (8 * L) ** 2 / 1e6
print(f"decode {8 * L}x{8 * L}: {min(ts) * 1e3:.1f} ms  {mp / min(ts):.1f} MP/s   out max {st['out_max']:.3f} hdr_pixels {st['hdr_pixels']} accepted {st['accepted']}")
eng.lib.hdrvae_profile_begin()
eng.decode(z, "aggressive", want_stats=False)
eng.lib.hdrvae_profile_end(f"gpurun_out/profile_big_{L}.tsv".encode())
if check:
    from oracle import hdr_oracle as ho
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    eng._workspace = None
    torch.cuda.empty_cache()
    dec = dec.to(dev)
    t0 = time.perf_counter()
    ref, rst, _ = ho.simple_hdr_decode(dec, z, "aggressive", 1.0)
    torch.cuda.synchronize()
    print(f"fp32 PyTorch eager oracle on the same GPU: {time.perf_counter() - t0:.1f} s")
    rel = float((out.double() - ref.double()).norm() / ref.double().norm())
    print(f"rel-L2 vs fp32 oracle: {rel:.3e}   (stats pre_max {st['pre_max']:.4f} vs {rst['pre_max']:.4f})")
