"""CPU fp32 restatement of the reference's HDR decode math (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/hdr_vae_decode.py; every function cites the lines it
restates.  The restatement uses the same PyTorch CPU elementwise ops the
reference uses, in the same order, so on one machine it reproduces the
reference bit for bit; it is pinned by tests/test_oracle_cpu.py against
tests/golden/*.npz, which hold outputs of the UNMODIFIED reference node run in
the build container (generator: oracle/make_golden.py).

Only the happy path is restated.  The bypass fallback
ladder (hdr_vae_decode.py:125-174, 443-835, 1205-1341) is non-deterministic
(fresh randomly initialised adapter convs) and cannot be pinned;
the deterministic replacement rule is documented in DESIGN.md and implemented
identically here and in the CUDA path:

  * no HDR data (max of the 3-channel max-pool <= 1.001): ``adaptive_recovery``
    and ``mathematical_recovery`` return the linearised LDR image (the
    reference's own "default result as a fallback", :1105); ``conservative`` and
    ``exposure`` are well defined in the reference and are restated as is;
  * accept test fails (no output value > 1.0 and max <= 1.1, :100-112): the
    intelligent result is returned anyway and ``stats['accepted']`` is 0.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

# hdr_vae_decode.py:48 — the enum as it is in code (default mathematical_recovery)
HDR_MODES = ("conservative", "exposure", "adaptive_recovery", "mathematical_recovery")
# README.md:37,78-81 names (BASELINE.json north_star) -> code behaviour
MODE_ALIASES = {
    "moderate": ("conservative", 3.0),          # smart_hdr_expansion(..., expansion_factor=3.0)
    "aggressive": ("mathematical_recovery", 1.0),
}

NORM_NONE, NORM_SIGMOID, NORM_TANH = 0, 1, 2


def resolve_mode(hdr_mode: str) -> Tuple[str, float]:
    """-> (code mode, smart-expansion factor).  Factor is 1.0 for the code names
    because simple_hdr_decode does not forward conservative_ev_multiplier
    (hdr_vae_decode.py:97 vs :1009,1107)."""
    m = hdr_mode.lower()
    if m in MODE_ALIASES:
        return MODE_ALIASES[m]
    if m not in HDR_MODES:
        raise ValueError(f"unknown hdr_mode {hdr_mode!r}")
    return m, 1.0


def channel_maxpool3(pre: torch.Tensor) -> torch.Tensor:
    """[B,128,H,W] -> [B,H,W,3]: max over channels 0-41 / 42-83 / 84-125,
    channels 126,127 ignored (hdr_vae_decode.py:1042-1056; same at :231-254)."""
    r, _ = torch.max(pre[:, 0:42], dim=1, keepdim=True)
    g, _ = torch.max(pre[:, 42:84], dim=1, keepdim=True)
    b, _ = torch.max(pre[:, 84:126], dim=1, keepdim=True)
    # same memory layout as the reference (NCHW storage viewed as BHWC): CPU vectorised
    # pow/log2 kernels round the loop tails differently for other layouts (1 ulp)
    return torch.cat([r, g, b], dim=1).permute(0, 2, 3, 1)


def channel_argmax3(pre: torch.Tensor) -> torch.Tensor:
    """First-max channel index of each pooled group (build-side extension used
    for the bit-exact integer check; torch.max(dim) returns the first max)."""
    r = pre[:, 0:42].argmax(dim=1)
    g = pre[:, 42:84].argmax(dim=1) + 42
    b = pre[:, 84:126].argmax(dim=1) + 84
    return torch.stack([r, g, b], dim=-1).to(torch.int32)


def srgb_to_linear(s: torch.Tensor) -> torch.Tensor:
    """hdr_vae_decode.py:1163-1203."""
    a = torch.abs(s)
    lin = torch.where(a <= 0.04045, a / 12.92, torch.pow((a + 0.055) / 1.055, 2.4))
    return torch.sign(s) * lin


def inverse_sigmoid(x: torch.Tensor) -> torch.Tensor:
    """hdr_vae_decode.py:927-932."""
    return torch.logit(torch.clamp(x, 1e-7, 1 - 1e-7))


def inverse_tanh(x: torch.Tensor) -> torch.Tensor:
    """hdr_vae_decode.py:934-939."""
    return torch.atanh(torch.clamp(x, -1 + 1e-6, 1 - 1e-6))


def analyze(pre: torch.Tensor, conv_w: torch.Tensor, conv_b: torch.Tensor, conv_fn=None) -> Dict:
    """analyze_conv_out (hdr_vae_decode.py:837-925) given the hooked tensor.

    ``pre`` is the input of decoder.conv_out ([B,128,H,W]); the "final result"
    is ComfyUI's VAE.decode output clamp((conv+1)/2,0,1) in BHWC."""
    # conv_fn: the same convolution evaluated band by band (oracle/big_oracle.py) for tensors of >= 2^31 elements
    conv_only = conv_fn(pre) if conv_fn is not None else F.conv2d(pre, conv_w, conv_b, padding=1)   # :876
    standard = torch.clamp((conv_only + 1.0) / 2.0, 0.0, 1.0).movedim(1, -1)  # comfy.sd.VAE.decode
    st = {
        "pre_min": float(pre.min()), "pre_max": float(pre.max()),             # :862-865
        "pre_mean": float(pre.mean()), "pre_std": float(pre.std()),
        "post_min": float(standard.min()), "post_max": float(standard.max()), # :867-870
        "post_mean": float(standard.mean()), "post_std": float(standard.std()),
        "conv_min": float(conv_only.min()), "conv_max": float(conv_only.max()),  # :877-879
        "conv_mean": float(conv_only.mean()),
    }
    if abs(st["post_max"] - 1.0) < 1e-3 and abs(st["post_min"] - 0.0) < 1e-3:    # :890-892
        st["norm_function"] = NORM_SIGMOID
    elif abs(st["post_max"] - 1.0) < 1e-3 and abs(st["post_min"] + 1.0) < 1e-3:  # :893-895
        st["norm_function"] = NORM_TANH
    else:
        st["norm_function"] = NORM_NONE                                        # :896-897 (fresh instance)
    return {"conv_only": conv_only, "standard": standard, "stats": st}


def intelligent(standard: torch.Tensor, pre: torch.Tensor, st: Dict, mode: str,
                expansion_factor: float = 1.0) -> Tuple[torch.Tensor, Dict]:
    """intelligent_hdr_decode (hdr_vae_decode.py:1009-1161) after the decode."""
    pre3 = channel_maxpool3(pre)                                               # :1042-1056
    out_st = {"pre3_min": float(pre3.min()), "pre3_max": float(pre3.max())}    # :1065-1066
    ldr = srgb_to_linear(standard)                                             # :1074
    has_hdr = out_st["pre3_max"] > (1.0 + 1e-3)                                # :1076-1078
    out_st["has_hdr"] = int(has_hdr)
    map_rec = pre3                                                             # :1080
    aligned = None                                                             # :1081 (python float 1.0)
    if has_hdr:
        if st["norm_function"] == NORM_TANH:                                   # :1085-1093
            rec = inverse_tanh(standard)
        elif st["norm_function"] == NORM_SIGMOID:
            rec = inverse_sigmoid(standard)
        else:
            rec = standard
        rmin, rmax = torch.min(rec), torch.max(rec)
        out_st["rec_min"], out_st["rec_max"] = float(rmin), float(rmax)
        original_range = st["pre_max"] - st["pre_min"]                         # :1097 (python doubles)
        rn = (rec - rmin) / (rmax - rmin)                                      # :1098
        map_rec = rn * original_range + st["pre_min"]                          # :1099
        aligned = map_rec - st["pre_mean"] + 1.0                               # :1102

    out_st["highlight_count"] = int(torch.sum(pre3 > 1.0))                     # :960-961
    if mode == "conservative":                                                 # :1106-1108, :941-980
        mask = pre3 > 1.0
        if out_st["highlight_count"] > 0:
            result = torch.where(mask, ldr + (pre3 - 1.0) * expansion_factor * ldr, ldr)
        else:
            result = ldr.clone()
    elif mode == "exposure":                                                   # :1110-1112, :982-1007
        result = ldr * torch.pow(2.0, torch.log2(torch.clamp(map_rec, min=0.001)))
    elif mode == "adaptive_recovery":                                          # :1114-1147
        if aligned is None:
            result = ldr                       # reference raises TypeError here -> bypass; see module doc
        else:
            amax = torch.max(aligned)
            out_st["aligned_max"] = float(amax)
            cf = 1.0
            if amax > 1.0 and amax > st["pre_max"]:
                cf = (st["pre_max"] - 1.0) / (amax - 1.0)     # python double / fp32 0-dim tensor -> fp32 tensor
            hm = (aligned > 1.0).float()
            mc = aligned * (1.0 - hm) + ((aligned - 1.0) * cf + 1.0) * hm
            result = ldr * torch.pow(2.0, torch.log2(torch.clamp(mc, min=0.001)))
    elif mode == "mathematical_recovery":                                      # :1149-1159
        if aligned is None:
            result = ldr
        else:
            result = ldr * torch.pow(2.0, torch.log2(torch.clamp(aligned, min=0.001)))
    else:
        raise ValueError(mode)
    return result, out_st


def hdr_epilogue(pre: torch.Tensor, conv_w: torch.Tensor, conv_b: torch.Tensor,
                 hdr_mode: str = "mathematical_recovery",
                 conservative_ev_multiplier: float = 1.0, conv_fn=None) -> Tuple[torch.Tensor, Dict]:
    """Everything simple_hdr_decode does after the decoder produced ``pre``
    (hdr_vae_decode.py:88-112, 180-195) -> (float32 [B,H,W,3] contiguous, stats)."""
    mode, factor = resolve_mode(hdr_mode)
    pre = pre.float()
    an = analyze(pre, conv_w.float(), conv_b.float(), conv_fn)
    st = dict(an["stats"])
    decoded, st2 = intelligent(an["standard"], pre, st, mode, factor)
    st.update(st2)
    hdr_pixels = int(torch.sum(decoded > 1.0))                                 # :100-102
    dmax = float(torch.max(decoded))
    st["accepted"] = int(hdr_pixels > 0 or dmax > 1.1)                         # :106
    if conservative_ev_multiplier != 1.0:                                      # :180-182
        decoded = decoded * conservative_ev_multiplier
    out = decoded.contiguous().float()                                         # :209-212, :354
    st["out_min"], st["out_max"] = float(out.min()), float(out.max())          # :188-189
    st["hdr_pixels"] = int(torch.sum(out > 1.0))                               # :190
    st["negative_pixels"] = int(torch.sum(out < 0.0))                          # :191
    return out, st


@torch.no_grad()
def simple_hdr_decode(decoder, latent: torch.Tensor, hdr_mode: str = "mathematical_recovery",
                      conservative_ev_multiplier: float = 1.0):
    """Whole node call (hdr_vae_decode.py:62-195) on an oracle FluxDecoder:
    ONE decoder pass (the reference runs two identical ones, :859 and :1022)."""
    p = next(decoder.parameters())
    pre = decoder.features(latent.to(p.device, p.dtype)).float()
    out, st = hdr_epilogue(pre, decoder.conv_out.weight, decoder.conv_out.bias,
                           hdr_mode, conservative_ev_multiplier)
    return out, st, pre
