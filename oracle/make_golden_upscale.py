"""Generate tests/golden/up_*.npz by running the UNMODIFIED reference upscaler node
(/root/reference/hdr_upscale_with_model.py:148-263) on CPU with the stand-ins of oracle.ref_loader.

TEST INFRASTRUCTURE ONLY; runs only where /root/reference is mounted.  Usage: python -m oracle.make_golden_upscale
Weights are re-created from the seed by oracle.upscaler_oracle.build_upscaler; a fingerprint guards RNG drift.
The restated `upscale()` must agree with the reference node bit for bit on every case (asserted here).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import upscaler_oracle as uo  # noqa: E402
from oracle.ref_loader import load_reference_upscaler  # noqa: E402

CASES = {
    # name: (B, H, W, image seed, image gain, net blocks, conv_last gain, small_blur, local_fix, method)
    "up_a_single_tile": (1, 24, 20, 11, 3.0, 2, 1.0, False, False, "bilinear"),
    "up_b_two_tiles": (1, 520, 16, 12, 3.0, 1, 1.0, False, False, "bilinear"),
    "up_c_saturating_blur_fix": (2, 20, 28, 13, 2.0, 2, 40.0, True, True, "bilinear"),
}


def make_image(b, h, w, seed, gain):
    """HDR-like test image: |N(0,1)| * gain / 3 with a few strong highlights (values well above 1)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(b, h, w, 3, generator=g)
    hot = torch.rand(b, h, w, 1, generator=g) > 0.9
    return torch.where(hot, img * gain * 2.0, img * 0.9)


def fingerprint(net):
    return float(sum(float(p.double().abs().sum()) for p in net.parameters()))


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, (b, h, w, seed, gain, nb, last_gain, blur, fix, method) in CASES.items():
        net = uo.build_upscaler(0, nb=nb, gain=last_gain)
        desc = uo.FakeDescriptor(net, 4, "ESRGAN")
        img = make_image(b, h, w, seed, gain)
        node = load_reference_upscaler(desc)
        with torch.no_grad():
            (ref,) = node.upscale(img, "fake_esrgan.pth", blur, fix, method)
        mine = uo.upscale(img, desc, blur, fix, method)
        assert ref.shape == (b, 4 * h, 4 * w, 3) and torch.equal(ref, mine), (name, float((ref - mine).abs().max()))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), image=img.numpy(), output=ref.numpy(),
                            params=np.array([b, h, w, seed, nb], dtype=np.int64), gain=np.array([gain, last_gain]),
                            flags=np.array([int(blur), int(fix)]), method=np.array(method),
                            weight_fingerprint=np.array(fingerprint(net)))
        print(f"{name}: out {tuple(ref.shape)} range [{float(ref.min()):.4f}, {float(ref.max()):.4f}] == restated oracle")


if __name__ == "__main__":
    main()
