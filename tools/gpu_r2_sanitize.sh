#!/bin/bash
# compute-sanitizer memcheck on the fused attention kernel, the fused-shortcut conv path and one small decode
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
cat > /tmp/san_case.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from vae_decode_hdr_b200.engine import HdrVaeEngine
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict, synthetic_latent
dev = torch.device("cuda:0")
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
for T in (24, 384, 1000):
    q = torch.randn(1, T, 512, device=dev).half(); k = torch.randn(1, T, 512, device=dev).half(); v = torch.randn(1, T, 512, device=dev).half()
    o = eng.attention(q, k, v); torch.cuda.synchronize(); print("attention", T, bool(torch.isfinite(o).all()))
for (b, h, w) in ((1, 8, 8), (2, 5, 9)):
    out, st = eng.decode(synthetic_latent(b, h, w).to(dev), "exposure"); torch.cuda.synchronize(); print("decode", b, h, w, st["hdr_pixels"])
    out, st = eng.decode(synthetic_latent(b, h, w).to(dev), "exposure"); torch.cuda.synchronize()      # second call: graph capture
    out, st = eng.decode(synthetic_latent(b, h, w).to(dev), "exposure"); torch.cuda.synchronize()      # third: replay
eng2 = HdrVaeEngine(random_decoder_state_dict(0), dev, precision="high")
out, st = eng2.decode(synthetic_latent(1, 8, 8).to(dev), "moderate"); torch.cuda.synchronize(); print("high", st["hdr_pixels"])
print("done")
PY
python /tmp/san_case.py > gpurun_out/r2_san_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_case.py > gpurun_out/r2_san_memcheck.log 2>&1
echo "memcheck exit $?"
tail -12 gpurun_out/r2_san_memcheck.log
