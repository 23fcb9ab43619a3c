"""Memory-lean fp32 evaluation of the oracle decoder for very large latents (TEST INFRASTRUCTURE ONLY).

`oracle.flux_decoder.FluxDecoder.features` is the oracle; at BASELINE config C4 (1x16x512x512 latent -> 4096x4096)
its widest activations hold 2^32 elements, beyond what single cuDNN / ATen calls index.  This module evaluates the SAME
graph with the SAME weights band by band:

  * every 3x3 / 1x1 convolution runs over horizontal bands with a one-row halo (zero rows outside the image), which
    is the same arithmetic as the whole-image convolution (reference call sites: hdr_vae_decode.py:859,:1022 ->
    ComfyUI Decoder; graph in SURVEY.md §8 a3);
  * GroupNorm (32 groups, eps 1e-6, biased variance) takes its statistics over the whole image in fp64, chunk by
    chunk, and is applied band by band in fp32;
  * nearest-2x upsampling is done band by band;
  * the single-head attention is query-chunked exactly like `Attn.forward` (softmax(q k^T / sqrt(c)) v per chunk of
    query rows against ALL keys).

`tests/test_oracle_cpu.py::test_banded_oracle_equals_plain_oracle` pins it to `FluxDecoder.features` at sizes both
can run.  Nothing in the product imports this file.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .flux_decoder import GN_EPS, GN_GROUPS


def conv_banded(x: torch.Tensor, conv: torch.nn.Conv2d, band: int = 256, out: torch.Tensor | None = None) -> torch.Tensor:
    """conv(x) for a stride-1 'same' conv (3x3 pad 1 or 1x1) evaluated over bands of `band` output rows."""
    B, C, H, W = x.shape
    k = conv.kernel_size[0]
    halo = k // 2
    if out is None:
        out = torch.empty((B, conv.out_channels, H, W), dtype=x.dtype, device=x.device)
    for y0 in range(0, H, band):
        y1 = min(H, y0 + band)
        a, b = max(0, y0 - halo), min(H, y1 + halo)
        xb = x[:, :, a:b]
        pt, pb = halo - (y0 - a), halo - (b - y1)          # zero rows outside the image
        if pt or pb:
            xb = F.pad(xb, (0, 0, pt, pb))
        out[:, :, y0:y1] = F.conv2d(xb, conv.weight, conv.bias, padding=(0, halo))
    return out


def groupnorm_banded(x: torch.Tensor, gn: torch.nn.GroupNorm, silu: bool, band: int = 256, inplace: bool = False) -> torch.Tensor:
    """[silu](GroupNorm(x)): statistics in fp64 over the whole image, applied band by band."""
    B, C, H, W = x.shape
    G = GN_GROUPS
    cpg = C // G
    s = torch.zeros((B, G), dtype=torch.float64, device=x.device)
    for y0 in range(0, H, band):
        s += x[:, :, y0:y0 + band].double().reshape(B, G, -1).sum(-1)
    n = float(cpg * H * W)
    mean = s / n
    m2 = torch.zeros_like(s)
    for y0 in range(0, H, band):                              # two-pass variance: no cancellation
        d = x[:, :, y0:y0 + band].double().reshape(B, G, -1) - mean[:, :, None]
        m2 += (d * d).sum(-1)
    rstd = 1.0 / torch.sqrt(m2 / n + GN_EPS)
    scale = (rstd[:, :, None] * gn.weight.double().reshape(1, G, cpg)).reshape(B, C, 1, 1)
    shift = (gn.bias.double().reshape(1, G, cpg) - mean[:, :, None] * rstd[:, :, None] * gn.weight.double().reshape(1, G, cpg)).reshape(B, C, 1, 1)
    scale, shift = scale.float(), shift.float()
    y = x if inplace else torch.empty_like(x)
    for y0 in range(0, H, band):
        t = x[:, :, y0:y0 + band] * scale + shift
        y[:, :, y0:y0 + band] = F.silu(t) if silu else t
    return y


def upsample2x_banded(x: torch.Tensor, band: int = 256) -> torch.Tensor:
    B, C, H, W = x.shape
    y = torch.empty((B, C, 2 * H, 2 * W), dtype=x.dtype, device=x.device)
    for y0 in range(0, H, band):
        y[:, :, 2 * y0:2 * min(H, y0 + band)] = F.interpolate(x[:, :, y0:y0 + band], scale_factor=2.0, mode="nearest")
    return y


def _res(blk, x, band):
    h = conv_banded(groupnorm_banded(x, blk.norm1, True, band), blk.conv1, band)
    h = groupnorm_banded(h, blk.norm2, True, band, inplace=True)
    h = conv_banded(h, blk.conv2, band)
    if hasattr(blk, "nin_shortcut"):
        x = conv_banded(x, blk.nin_shortcut, band)
    h += x
    return h


def _attn(at, x, q_chunk):
    b, c, h, w = x.shape
    T = h * w
    hn = groupnorm_banded(x, at.norm, False)
    q = at.q(hn).reshape(b, c, T).transpose(1, 2)
    k = at.k(hn).reshape(b, c, T)
    v = at.v(hn).reshape(b, c, T).transpose(1, 2)
    out = torch.empty_like(q)
    scale = 1.0 / math.sqrt(c)
    for s in range(0, T, q_chunk):
        p = torch.softmax(torch.bmm(q[:, s:s + q_chunk], k) * scale, dim=-1)
        out[:, s:s + q_chunk] = torch.bmm(p, v)
    out = out.transpose(1, 2).reshape(b, c, h, w)
    return x + at.proj_out(out)


@torch.no_grad()
def features_banded(dec, z: torch.Tensor, band: int = 256, q_chunk: int = 512) -> torch.Tensor:
    """== dec.features(z): SiLU(norm_out(.)) of the Flux AE decoder, [B,128,8h,8w] fp32."""
    h = conv_banded(z, dec.conv_in, band)
    h = _res(dec.mid.block_1, h, band)
    h = _attn(dec.mid.attn_1, h, q_chunk)
    h = _res(dec.mid.block_2, h, band)
    for lvl in reversed(range(len(dec.up))):
        for blk in dec.up[lvl].block:
            h = _res(blk, h, band)
        if lvl != 0:
            h = conv_banded(upsample2x_banded(h, band), dec.up[lvl].upsample.conv, band)
    return groupnorm_banded(h, dec.norm_out, True, band, inplace=True)


@torch.no_grad()
def simple_hdr_decode_banded(dec, latent: torch.Tensor, hdr_mode: str = "mathematical_recovery",
                             conservative_ev_multiplier: float = 1.0, band: int = 256, q_chunk: int = 512):
    """== hdr_oracle.simple_hdr_decode for latents whose activations exceed 2^31 elements: banded decoder, banded
    conv_out, the HDR math itself unchanged."""
    from . import hdr_oracle as ho
    p = next(dec.parameters())
    pre = features_banded(dec, latent.to(p.device, p.dtype), band, q_chunk).float()
    out, st = ho.hdr_epilogue(pre, dec.conv_out.weight, dec.conv_out.bias, hdr_mode, conservative_ev_multiplier,
                              conv_fn=lambda t: conv_banded(t, dec.conv_out, band))
    return out, st, pre
