"""Kernel-level parity on the B200, through the C ABI (vae_decode_hdr_b200.engine -> libhdrvae.so).
Each CUDA kernel against a plain PyTorch fp32 evaluation of the same op on the same bf16-rounded
operands (tolerances written at each check)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def engine():
    from oracle.flux_decoder import build_decoder
    from vae_decode_hdr_b200.engine import HdrVaeEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dec = build_decoder(0)
    eng = HdrVaeEngine(dec.state_dict(), DEV)
    yield eng
    eng.close()


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _bf(x):
    return x.to(torch.bfloat16).float()


def _conv_ref(x_nhwc, w, b, ksize, upsample, residual):
    """fp32 conv on the bf16-rounded activations, fp32 weights -> NHWC."""
    x = x_nhwc.float().permute(0, 3, 1, 2)
    if upsample:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    y = F.conv2d(x, w, b, padding=ksize // 2).permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    return y


CONV_CASES = [
    # B, H, W, Cin, Cout, ksize, upsample, residual
    (1, 8, 16, 64, 128, 3, False, False),     # exactly one 128-pixel tile, one k-block per tap
    (1, 16, 16, 128, 128, 3, False, True),
    (2, 12, 20, 128, 256, 3, False, True),    # ragged tiles (masked rows), BLOCK_N = 256
    (1, 32, 32, 512, 512, 3, False, True),    # two n-tiles, K = 4608
    (1, 16, 16, 256, 128, 1, False, False),   # nin_shortcut shape
    (1, 8, 8, 256, 256, 3, True, False),      # upsample folded into the load
    (2, 5, 7, 128, 128, 3, True, False),      # odd sizes + upsample
    (1, 3, 5, 64, 32, 3, False, False),       # smaller than one tile, narrow N
    (3, 64, 64, 128, 128, 3, False, False),   # many tiles -> several per CTA (pipeline phase wrap)
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
def test_conv_tcgen05_matches_torch_and_direct(engine, case):
    from vae_decode_hdr_b200 import _native as N
    B, H, W, cin, cout, ks, up, res = case
    g = torch.Generator(device="cpu").manual_seed(hash(case) % (2 ** 31))
    x = _bf(torch.randn(B, H, W, cin, generator=g)).to(DEV)
    w = (torch.randn(cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks)).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    OH, OW = (2 * H, 2 * W) if up else (H, W)
    r = _bf(torch.randn(B, OH, OW, cout, generator=g)).to(DEV) if res else None
    y_tc = engine.conv2d(x.bfloat16(), w, b, ks, up, r.bfloat16() if res else None, out_f32=True, impl=N.CONV_TCGEN05)
    y_dr = engine.conv2d(x.bfloat16(), w, b, ks, up, r.bfloat16() if res else None, out_f32=True, impl=N.CONV_DIRECT)
    torch.cuda.synchronize()
    assert torch.isfinite(y_tc).all()
    # same bf16 operands, fp32 accumulation in a different order: 1e-5 relative
    assert _rel(y_tc, y_dr) < 1e-5, ("tcgen05 vs direct", _rel(y_tc, y_dr))
    # torch fp32 conv with UNROUNDED weights: the bf16 weight rounding (2^-9 relative per weight,
    # averaged over K terms) bounds the difference
    ref = _conv_ref(x, w, b, ks, up, r)
    assert _rel(y_tc, ref) < 4e-3, ("tcgen05 vs torch", _rel(y_tc, ref))
    # bf16 output path
    y_bf = engine.conv2d(x.bfloat16(), w, b, ks, up, r.bfloat16() if res else None, out_f32=False)
    assert _rel(y_bf.float(), y_tc) < 4e-3


def test_conv_weight_rounding_is_the_only_difference(engine):
    """With weights that are exactly representable in bf16 the kernel must agree with torch to fp32 noise."""
    g = torch.Generator().manual_seed(5)
    x = _bf(torch.randn(1, 16, 24, 128, generator=g)).to(DEV)
    w = _bf(torch.randn(256, 128, 3, 3, generator=g) / 34.0).to(DEV)
    b = torch.randn(256, generator=g).to(DEV)
    y = engine.conv2d(x.bfloat16(), w, b, 3, out_f32=True)
    ref = _conv_ref(x, w, b, 3, False, None)
    assert _rel(y, ref) < 2e-6


@pytest.mark.parametrize("shape,silu", [((2, 16, 16, 512), True), ((1, 9, 13, 256), True), ((3, 32, 32, 128), True),
                                        ((1, 64, 64, 512), False), ((1, 1, 1, 128), True)])
def test_groupnorm_silu(engine, shape, silu):
    g = torch.Generator().manual_seed(11)
    B, H, W, Cc = shape
    x = _bf(torch.randn(shape, generator=g) * 2.0 + 0.7).to(DEV)
    gamma = (1.0 + 0.1 * torch.randn(Cc, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(Cc, generator=g)).to(DEV)
    y = engine.groupnorm_silu(x.bfloat16(), gamma, beta, silu).float()
    ref = F.group_norm(x.permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-6)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    # output is rounded to bf16: half an ulp = 2^-9 relative
    assert torch.allclose(y, ref, rtol=2 ** -8, atol=2e-3), float((y - ref).abs().max())
    assert _rel(y, ref) < 3e-3


def test_groupnorm_is_run_to_run_deterministic(engine):
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 40, 40, 256, generator=g).to(DEV).bfloat16()
    gamma = torch.ones(256, device=DEV)
    beta = torch.zeros(256, device=DEV)
    a = engine.groupnorm_silu(x, gamma, beta, True)
    b = engine.groupnorm_silu(x, gamma, beta, True)
    assert torch.equal(a, b)


@pytest.mark.parametrize("B,T", [(1, 64), (2, 256), (1, 24), (1, 1000), (1, 4096)])
def test_attention(engine, B, T):
    g = torch.Generator().manual_seed(T)
    q = _bf(torch.randn(B, T, 512, generator=g)).to(DEV)
    k = _bf(torch.randn(B, T, 512, generator=g)).to(DEV)
    v = _bf(torch.randn(B, T, 512, generator=g)).to(DEV)
    o = engine.attention(q.bfloat16(), k.bfloat16(), v.bfloat16()).float()
    p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(512.0), dim=-1)
    ref = p @ v
    # probabilities are rounded to bf16 before PV and the output to bf16: ~2^-9 relative each
    assert _rel(o, ref) < 6e-3, _rel(o, ref)


def test_pack_half_bit_exact_vs_numpy():
    from vae_decode_hdr_b200.engine import pack_half
    g = torch.Generator().manual_seed(3)
    img = torch.randn(2, 7, 9, 3, generator=g) * 100.0
    img[0, 0, 0, 0] = 70000.0       # overflow -> inf (linear_exr_export.py:155 numpy astype semantics)
    img[0, 0, 0, 1] = -1e-8         # underflow -> -0
    img[0, 0, 0, 2] = 6.0e-6        # fp16 subnormal
    img[0, 0, 1, 0] = 1.00048828125  # exact tie -> round to even
    want = img.numpy().astype(np.float16)
    got = pack_half(img.to(DEV)).cpu().numpy()
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16))
    planes = pack_half(img.to(DEV), exr_scanline_order=True).cpu().numpy()     # [B,H,3(B,G,R),W]
    assert np.array_equal(planes.view(np.uint16), np.ascontiguousarray(want[..., ::-1].transpose(0, 1, 3, 2)).view(np.uint16))
    assert pack_half(torch.empty(0, 4, 4, 3, device=DEV)).shape == (0, 4, 4, 3)
