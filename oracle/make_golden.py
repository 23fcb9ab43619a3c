"""Generate tests/golden/*.npz by running the UNMODIFIED reference node.

TEST INFRASTRUCTURE ONLY; runs only where /root/reference is mounted (the build
container).  Usage:  python -m oracle.make_golden

For every case the script records
  * the latent and the decoder variant (weights are re-created from the seed by
    oracle.flux_decoder.build_decoder; a fingerprint guards against RNG drift),
  * ``pre_conv_out`` as captured by the reference's own forward hook
    (hdr_vae_decode.py:850-855) plus conv_out weight/bias,
  * the reference's analysis stats (hdr_vae_decode.py:912-919),
  * ``intelligent_hdr_decode`` output for each mode (hdr_vae_decode.py:1009),
  * the full node output ``simple_hdr_decode`` (hdr_vae_decode.py:62) whenever
    the reference accepted the intelligent result (no bypass).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.flux_decoder import FakeComfyVAE, build_decoder, make_latent, weight_fingerprint  # noqa: E402
from oracle.ref_loader import load_reference_node  # noqa: E402

MODES = ["conservative", "exposure", "adaptive_recovery", "mathematical_recovery"]

CASES = {
    # name: (batch, h, w, latent_seed, variant, ev multipliers to record for the node call)
    "a_b1_4x4": (1, 4, 4, 1234, "default", [1.0, 2.5]),
    "b_b2_4x6": (2, 4, 6, 7, "default", [1.0]),
    "c_nohdr_b1_4x4": (1, 4, 4, 1234, "nohdr", []),
    "d_nonorm_b1_4x4": (1, 4, 4, 1234, "nonorm", []),
}


def apply_variant(dec, variant: str):
    with torch.no_grad():
        if variant == "nohdr":        # nothing in pre_conv_out exceeds 1.0 (SURVEY.md §0.8 probe)
            dec.norm_out.weight.mul_(0.05)
        elif variant == "nonorm":     # conv_out never saturates -> neither SIGMOID nor TANH detected
            dec.conv_out.weight.mul_(0.1)
            dec.conv_out.bias.mul_(0.1)
        elif variant != "default":
            raise ValueError(variant)
    return dec


def run_case(name, b, h, w, seed, variant, mults):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    dec = apply_variant(build_decoder(0), variant)
    vae = FakeComfyVAE(dec)
    z = make_latent(b, h, w, seed)
    rec = {
        "latent": z.numpy(), "variant": np.array(variant), "latent_seed": np.array(seed),
        "weight_fingerprint": np.array(weight_fingerprint(dec)),
        "conv_w": dec.conv_out.weight.detach().numpy().copy(),
        "conv_b": dec.conv_out.bias.detach().numpy().copy(),
    }
    node = load_reference_node()
    analysis = node.analyze_conv_out(vae, z)
    rec["pre_conv_out"] = analysis["pre_conv_out"].numpy().copy()
    rec["final_result"] = analysis["final_result"].numpy().copy()
    rec["norm_function"] = np.array(node.NORMALIZATION_FUNCTION)
    for grp in ("pre_stats", "post_stats", "conv_stats"):
        for k, v in analysis[grp].items():
            rec[f"{grp}.{k}"] = np.array(v, dtype=np.float64)
    for mode in MODES:
        try:
            out = node.intelligent_hdr_decode(vae, z, analysis, mode)
            rec[f"intelligent.{mode}"] = out.numpy().copy()
        except Exception as e:  # adaptive/mathematical raise TypeError without HDR data
            rec[f"intelligent.{mode}.error"] = np.array(type(e).__name__)
    for mult in mults:
        for mode in MODES:
            fresh = load_reference_node()
            (img,) = fresh.simple_hdr_decode({"samples": z}, vae, hdr_mode=mode,
                                             conservative_ev_multiplier=mult)
            rec[f"node.{mode}.x{mult}"] = img.numpy().copy()
    out_path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(out_path, **rec)
    print(name, {k: (v.shape if v.ndim else v.item()) for k, v in rec.items()
                 if not k.startswith(("pre_conv", "latent", "conv_w"))})


def main():
    import logging
    logging.disable(logging.CRITICAL)
    for name, args in CASES.items():
        run_case(name, *args)


if __name__ == "__main__":
    main()
