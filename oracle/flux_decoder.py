"""fp32 PyTorch restatement of the Flux.1 AE *decoder* (TEST INFRASTRUCTURE ONLY).

The reference never implements the decoder: it calls ``vae.decode(latent)`` and
reaches into ``vae.first_stage_model.decoder`` (reference
hdr_vae_decode.py:842,855,859,876,1022).  The arithmetic is ComfyUI's
``comfy.ldm.modules.diffusionmodules.model.Decoder`` (third party, un-vendored,
un-pinned: requirements.txt:2 "provided by ComfyUI").  This file restates the
published BFL Flux.1 autoencoder decoder graph with the same state-dict keys
ComfyUI / BFL checkpoints use, so weights can be exchanged by ``state_dict``:

    conv_in(16->512, 3x3) -> mid.block_1 -> mid.attn_1 -> mid.block_2
    -> up[3]: 3 x Res(512), upsample -> up[2]: 3 x Res(512), upsample
    -> up[1]: Res(512->256), 2 x Res(256), upsample
    -> up[0]: Res(256->128), 2 x Res(128)
    -> norm_out -> SiLU -> conv_out(128->3, 3x3)

GroupNorm: 32 groups, eps 1e-6, affine.  Upsample: nearest x2 then 3x3 conv.
Attention: single head, d = 512, scale 1/sqrt(512), over h*w tokens.

``FakeComfyVAE`` mimics the slice of ``comfy.sd.VAE`` the reference touches:
``.first_stage_model.decoder`` and ``.decode(z)`` =
``clamp((decoder(z)+1)/2, 0, 1).movedim(1, -1)``.

Parity: unpinned by the reference (no tests exist); cross-checked in
tests/test_oracle_cpu.py against the independent implementation shipped in this
image (torchtitan.experiments.flux.model.autoencoder.Decoder) when importable.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

CH = 128
CH_MULT = (1, 2, 4, 4)
NUM_RES_BLOCKS = 2
Z_CHANNELS = 16
OUT_CH = 3
GN_GROUPS = 32
GN_EPS = 1e-6


def _gn(c: int) -> nn.GroupNorm:
    return nn.GroupNorm(GN_GROUPS, c, eps=GN_EPS, affine=True)


class Res(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1 = _gn(cin)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.norm2 = _gn(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1, 1, 0)

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if hasattr(self, "nin_shortcut"):
            x = self.nin_shortcut(x)
        return x + h


class Attn(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.norm = _gn(c)
        self.q = nn.Conv2d(c, c, 1)
        self.k = nn.Conv2d(c, c, 1)
        self.v = nn.Conv2d(c, c, 1)
        self.proj_out = nn.Conv2d(c, c, 1)

    def forward(self, x):
        b, c, h, w = x.shape
        hn = self.norm(x)
        q = self.q(hn).reshape(b, c, h * w).transpose(1, 2)   # [b, T, c]
        k = self.k(hn).reshape(b, c, h * w)                   # [b, c, T]
        v = self.v(hn).reshape(b, c, h * w).transpose(1, 2)   # [b, T, c]
        # explicit softmax(QK^T/sqrt(c))V, row-chunked so T = 16k fits in RAM
        out = torch.empty_like(q)
        scale = 1.0 / math.sqrt(c)
        step = max(1, min(h * w, (1 << 26) // max(1, h * w)))
        for s in range(0, h * w, step):
            p = torch.softmax(torch.bmm(q[:, s:s + step], k) * scale, dim=-1)
            out[:, s:s + step] = torch.bmm(p, v)
        out = out.transpose(1, 2).reshape(b, c, h, w)
        return x + self.proj_out(out)


class Up(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 1, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class FluxDecoder(nn.Module):
    """Decoder with BFL/ComfyUI-compatible state-dict keys."""

    def __init__(self):
        super().__init__()
        top = CH * CH_MULT[-1]
        self.conv_in = nn.Conv2d(Z_CHANNELS, top, 3, 1, 1)
        self.mid = nn.Module()
        self.mid.block_1 = Res(top, top)
        self.mid.attn_1 = Attn(top)
        self.mid.block_2 = Res(top, top)
        self.up = nn.ModuleList()
        cin = top
        levels = []
        for lvl in reversed(range(len(CH_MULT))):
            cout = CH * CH_MULT[lvl]
            stage = nn.Module()
            stage.block = nn.ModuleList()
            stage.attn = nn.ModuleList()
            for _ in range(NUM_RES_BLOCKS + 1):
                stage.block.append(Res(cin, cout))
                cin = cout
            if lvl != 0:
                stage.upsample = Up(cin)
            levels.insert(0, stage)
        for s in levels:
            self.up.append(s)
        self.norm_out = _gn(cin)
        self.conv_out = nn.Conv2d(cin, OUT_CH, 3, 1, 1)

    def features(self, z):
        """Everything up to and including SiLU(norm_out(.)) = the tensor the
        reference's forward hook captures (hdr_vae_decode.py:850-855)."""
        h = self.conv_in(z)
        h = self.mid.block_1(h)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h)
        for lvl in reversed(range(len(self.up))):
            for blk in self.up[lvl].block:
                h = blk(h)
            if lvl != 0:
                h = self.up[lvl].upsample(h)
        return F.silu(self.norm_out(h))

    def forward(self, z):
        return self.conv_out(self.features(z))


def build_decoder(seed: int = 0, dtype=torch.float32, device="cpu") -> FluxDecoder:
    """Random-init decoder, PyTorch default init under ``torch.manual_seed(seed)``
    (SURVEY.md §8d "Synthetic inputs").  Always initialised on CPU so the
    weights are identical on every box, then moved."""
    gen_state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        dec = FluxDecoder()
    finally:
        torch.random.set_rng_state(gen_state)
    return dec.to(device=device, dtype=dtype).eval()


def make_latent(b: int, h: int, w: int, seed: int = 1234, device="cpu") -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(b, Z_CHANNELS, h, w, generator=g, dtype=torch.float32).to(device)


class _FirstStage:
    def __init__(self, decoder):
        self.decoder = decoder


class FakeComfyVAE:
    """The slice of comfy.sd.VAE the reference uses (hdr_vae_decode.py:842,859,1022)."""

    def __init__(self, decoder: FluxDecoder):
        self.first_stage_model = _FirstStage(decoder)
        p = next(decoder.parameters())
        self.device = p.device
        self.vae_dtype = p.dtype
        self.output_device = p.device

    @torch.no_grad()
    def decode(self, samples_in: torch.Tensor) -> torch.Tensor:
        dec = self.first_stage_model.decoder
        x = dec(samples_in.to(self.device).to(self.vae_dtype)).float()
        return torch.clamp((x + 1.0) / 2.0, 0.0, 1.0).to(self.output_device).movedim(1, -1)


def weight_fingerprint(dec: nn.Module) -> float:
    """Cheap checksum used by the goldens to detect an init/RNG drift."""
    s = 0.0
    for k, v in dec.state_dict().items():
        s += float(v.double().abs().sum())
    return s
