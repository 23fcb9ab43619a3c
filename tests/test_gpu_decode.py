"""End-to-end parity of the B200 decode against the fp32 oracle (same seeded weights and latents)."""
import numpy as np
import pytest
import torch

from oracle import hdr_oracle as ho
from oracle.flux_decoder import FakeComfyVAE, build_decoder, make_latent

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def setup():
    from vae_decode_hdr_b200.engine import HdrVaeEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dec = build_decoder(0).to(DEV)
    eng = HdrVaeEngine(dec.state_dict(), DEV)
    yield dec, eng
    eng.close()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("B,h,w", [(1, 8, 8), (2, 4, 6), (1, 16, 16), (1, 5, 9)])
def test_decoder_features_vs_oracle(setup, B, h, w):
    """SiLU(norm_out(h)) — the tensor the reference's hook captures — bf16 pipeline vs fp32 oracle.
    Tolerance: BASELINE.json north_star, rel-L2 <= 1e-2 for the bf16 decode."""
    dec, eng = setup
    z = make_latent(B, h, w, seed=100 + h * w).to(DEV)
    with torch.no_grad():
        ref = dec.features(z).permute(0, 2, 3, 1)
    got = eng.decode_features(z).float()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert _rel(got, ref) < 1e-2, _rel(got, ref)


def test_tcgen05_path_equals_direct_path(setup):
    """Whole decoder with the tcgen05 kernels vs the CUDA-core validation kernels: same bf16 operands,
    only the fp32 accumulation order differs (amplified a little by ~60 layers of bf16 re-rounding)."""
    from vae_decode_hdr_b200 import _native as N
    dec, eng = setup
    z = make_latent(1, 8, 8, seed=5).to(DEV)
    a = eng.decode_features(z).float()
    eng.set_conv_impl(N.CONV_DIRECT)
    try:
        b = eng.decode_features(z).float()
    finally:
        eng.set_conv_impl(N.CONV_TCGEN05)
    assert _rel(a, b) < 3e-3, _rel(a, b)


@pytest.mark.parametrize("mode", list(ho.HDR_MODES) + ["moderate"])
def test_full_decode_vs_oracle(setup, mode):
    """Node-level output.  bf16 decode within rel-L2 <= 1e-2 of the fp32 reference output (north_star)."""
    dec, eng = setup
    z = make_latent(2, 16, 16, seed=77).to(DEV)
    out, st = eng.decode(z, mode, 1.0)
    ref, rst, _ = ho.simple_hdr_decode(dec, z, mode, 1.0)
    assert out.shape == (2, 128, 128, 3) and out.dtype == torch.float32 and out.is_contiguous()
    assert st["accepted"] == rst["accepted"] == 1
    assert st["norm_function"] == rst["norm_function"] and st["has_hdr"] == rst["has_hdr"]
    assert _rel(out, ref.to(DEV)) < 1e-2, _rel(out, ref.to(DEV))
    assert st["pre_max"] == pytest.approx(rst["pre_max"], rel=2e-2)
    assert st["hdr_pixels"] == int((out > 1.0).sum())


def test_node_api_matches_reference_surface(setup):
    """Drop-in check: same INPUT_TYPES / RETURN_TYPES / FUNCTION / CATEGORY, same call, same output contract."""
    from vae_decode_hdr_b200 import NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS
    dec, _ = setup
    node_cls = NODE_CLASS_MAPPINGS["HDRVAEDecode"]
    assert NODE_DISPLAY_NAME_MAPPINGS["HDRVAEDecode"] == "HDR VAE Decode"
    it = node_cls.INPUT_TYPES()
    assert it["required"] == {"samples": ("LATENT",), "vae": ("VAE",)}
    assert it["optional"]["hdr_mode"][0] == ["conservative", "exposure", "adaptive_recovery", "mathematical_recovery"]
    assert it["optional"]["hdr_mode"][1]["default"] == "mathematical_recovery"
    assert it["optional"]["conservative_ev_multiplier"] == ("FLOAT", {"default": 1.0, "min": 0.1, "max": 10.0, "step": 0.1,
                                                            "tooltip": "Expansion multiplier for the conservative mode."})
    assert (node_cls.RETURN_TYPES, node_cls.RETURN_NAMES, node_cls.FUNCTION, node_cls.CATEGORY) == \
        (("IMAGE",), ("image",), "simple_hdr_decode", "latent")
    vae = FakeComfyVAE(dec)
    z_cpu = make_latent(1, 8, 8, seed=9)                      # ComfyUI hands CPU latents
    node = node_cls()
    (img,) = getattr(node, node_cls.FUNCTION)(samples={"samples": z_cpu}, vae=vae)      # defaults, like the executor
    assert img.dtype == torch.float32 and img.shape == (1, 64, 64, 3) and img.is_contiguous()
    assert img.device == vae.output_device
    ref, _, _ = ho.simple_hdr_decode(dec, z_cpu.to(DEV), "mathematical_recovery", 1.0)
    assert _rel(img, ref.to(img.device)) < 1e-2
    (img2,) = node.simple_hdr_decode({"samples": z_cpu}, vae, hdr_mode="conservative", conservative_ev_multiplier=2.0)
    (img1,) = node.simple_hdr_decode({"samples": z_cpu}, vae, hdr_mode="conservative", conservative_ev_multiplier=1.0)
    assert torch.allclose(img2, img1 * 2.0)                  # multiplier applies to every mode's final image (:180-182)
    assert node.NORMALIZATION_FUNCTION == "SIGMOID"
    with pytest.raises(ValueError):
        node.simple_hdr_decode({"samples": torch.zeros(0, 16, 8, 8)}, vae)
    with pytest.raises(ValueError):
        node.simple_hdr_decode({"samples": z_cpu}, vae, hdr_mode="nope")


def test_decode_is_deterministic(setup):
    _, eng = setup
    z = make_latent(1, 8, 8, seed=3).to(DEV)
    a, _ = eng.decode(z, "exposure")
    b, _ = eng.decode(z, "exposure")
    assert torch.equal(a, b)
