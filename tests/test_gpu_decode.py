"""End-to-end parity of the B200 decode against the fp32 oracle (same seeded weights and latents)."""
import pytest
import torch

from oracle import hdr_oracle as ho
from oracle.flux_decoder import FakeComfyVAE, build_decoder, make_latent

from _metrics import rel_l2_outside, saturation_band_mask

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def setup():
    from vae_decode_hdr_b200.engine import HdrVaeEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dec = build_decoder(0).to(DEV)
    eng = HdrVaeEngine(dec.state_dict(), DEV)          # default precision: fp16 operands, fp32 streams
    yield dec, eng
    eng.close()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("B,h,w", [(1, 8, 8), (2, 4, 6), (1, 16, 16), (1, 5, 9), (1, 32, 32)])
def test_decoder_features_vs_oracle(setup, B, h, w):
    """SiLU(norm_out(h)) — the tensor the reference's hook captures — 16-bit tensor-core pipeline vs the fp32
    oracle.  The end-to-end tolerance is rel-L2 <= 1e-2 on the IMAGE (north_star); the HDR math amplifies
    feature error about 4x (exposure mode), so the features are held to 3e-3 here (measured ~1.7e-3)."""
    dec, eng = setup
    z = make_latent(B, h, w, seed=100 + h * w).to(DEV)
    with torch.no_grad():
        ref = dec.features(z).permute(0, 2, 3, 1)
    got = eng.decode_features(z).float()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert _rel(got, ref) < 3e-3, _rel(got, ref)


def test_tcgen05_path_equals_direct_path(setup):
    """Whole decoder with the tcgen05 kernels vs the CUDA-core validation kernels: same rounded operands,
    only the fp32 accumulation order (and where the GroupNorm statistics are summed) differs.  Those last-bit
    differences flip fp16 roundings of the ~60 intermediate tensors, so two runs of the same 16-bit pipeline
    decorrelate to about sqrt(2) x the pipeline's own rounding error (~1.7e-3): bound 3e-3.  The tight
    tcgen05-vs-direct checks (1e-5) are per kernel in test_gpu_kernels.py."""
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
    """Node-level output within rel-L2 <= 1e-2 of the fp32 reference output (BASELINE.json north_star), plain rel-L2.
    BASELINE's own shapes (C1, C3, C4) and the known ill-conditioned case are in test_gpu_parity_big.py."""
    dec, eng = setup
    z = make_latent(2, 16, 16, seed=77).to(DEV)
    out, st = eng.decode(z, mode, 1.0)
    ref, rst, _ = ho.simple_hdr_decode(dec, z, mode, 1.0)
    assert out.shape == (2, 128, 128, 3) and out.dtype == torch.float32 and out.is_contiguous()
    assert st["accepted"] == rst["accepted"] == 1
    assert st["norm_function"] == rst["norm_function"] and st["has_hdr"] == rst["has_hdr"]
    assert _rel(out, ref.to(DEV)) < 1e-2, _rel(out, ref.to(DEV))
    assert st["pre_max"] == pytest.approx(rst["pre_max"], rel=5e-3)
    assert st["hdr_pixels"] == int((out > 1.0).sum())


def test_bf16_operand_mode_documented_error(setup):
    """HDRVAE_PRECISION_BF16: same kernels and speed, bf16 operand rounding (2^-8 instead of 2^-11).  It does
    NOT meet the 1e-2 image tolerance with random-init weights (DESIGN.md 'Precision'); bounds documented here."""
    from vae_decode_hdr_b200.engine import HdrVaeEngine
    dec, _ = setup
    eng = HdrVaeEngine(dec.state_dict(), DEV, precision="bf16")
    try:
        z = make_latent(1, 16, 16, seed=1234).to(DEV)
        with torch.no_grad():
            ref = dec.features(z).permute(0, 2, 3, 1)
        got = eng.decode_features(z)
        assert got.dtype == torch.bfloat16
        assert _rel(got.float(), ref) < 2e-2
        out, st = eng.decode(z, "conservative")
        oref, _, _ = ho.simple_hdr_decode(dec, z, "conservative", 1.0)
        assert _rel(out, oref.to(DEV)) < 2e-2
    finally:
        eng.close()


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


def test_batch_sharded_decode_equals_whole_batch(setup):
    """Multi-GPU batch sharding emulated on one GPU: two 'ranks' decode half the batch each, the raw statistics
    are merged exactly as the NCCL all-reduce does (MIN/MAX/SUM), and the result must equal the whole-batch
    decode up to the summation order of the double-precision sums (statistics are batch-global in the reference)."""
    from vae_decode_hdr_b200.sharding import merge_raw_stats, shard_bounds
    _, eng = setup
    z = make_latent(4, 8, 8, seed=31).to(DEV)
    whole, st = eng.decode(z, "adaptive_recovery")
    blocks, outs = [], []
    bounds = shard_bounds(4, 2)
    for s, e in bounds:
        vmin, vmax, vsum = eng.decode_begin(z[s:e])
        blocks.append((vmin.clone(), vmax.clone(), vsum.clone()))
    mmin, mmax, msum = merge_raw_stats(blocks)
    for s, e in bounds:
        vmin, vmax, vsum = eng.decode_begin(z[s:e])           # recompute the slice, then install the merged block
        vmin.copy_(mmin); vmax.copy_(mmax); vsum.copy_(msum)
        o, _ = eng.decode_finish("adaptive_recovery")
        outs.append(o)
    assert torch.allclose(torch.cat(outs), whole, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("world,h,w,mode", [(2, 8, 8, "mathematical_recovery"), (4, 8, 12, "exposure"), (2, 6, 5, "adaptive_recovery"),
                                            (4, 16, 16, "conservative"), (2, 16, 16, "moderate"), (3, 12, 16, "aggressive"),
                                            (2, 32, 8, "moderate")])
def test_row_tiled_decode_equals_single_gpu(setup, world, h, w, mode):
    """Spatial row tiling (config C4 path) with `world` virtual ranks on one GPU: conv halos exchanged, GroupNorm
    sums all-reduced, attention K/V all-gathered, HDR statistics all-reduced.  Tiling must not change the result:
      * the 1e-2 bar against the fp32 oracle holds for the tiled decode;
      * against the single-GPU decode only the fp32 summation order of the GroupNorm partials differs (tile
        shapes follow the slab height); that flips a few fp16 roundings which then decorrelate through the 30
        layers: measured 1.3e-3 on the image, bound 5e-3;
      * when the slabs tile exactly like the whole image (32x8 latent on 2 ranks: 16 latent rows per rank, so the
        8 x 16-pixel conv tiles and the 128-pixel row tiles fall on the same pixels at every level) every partial
        sum is identical and so is the image: any halo / gather / reduction slip shows up as a non-zero
        difference here."""
    from vae_decode_hdr_b200.sharding import decode_rows_emulated
    dec, eng = setup
    z = make_latent(1, h, w, seed=41 + world).to(DEV)
    tiled, st = decode_rows_emulated(eng, z, world, mode)
    whole, st1 = eng.decode(z, mode)
    assert tiled.shape == whole.shape
    ref, _, pre = ho.simple_hdr_decode(dec, z, mode, 1.0)
    # The logit-recovery modes are ill-conditioned at pixels whose sigmoid-domain value sits within 1e-3 of a clamp end
    # (logit slope > 1e3; the reference clamps at 1e-7): those pixels are identified BY RULE from the reference's own
    # conv_out values (tests/_metrics.py: saturation_band_mask) and set aside for the three logit modes; they must be a
    # small minority.  conservative / smart expansion are compared in full.
    if mode in ("conservative", "moderate"):
        band = torch.zeros(tiled.shape[:3], dtype=torch.bool, device=DEV)
    else:
        band = saturation_band_mask(dec, pre)
        assert float(band.float().mean()) < 1e-2
    assert rel_l2_outside(tiled, ref.to(DEV), band) < 1e-2, rel_l2_outside(tiled, ref.to(DEV), band)
    assert rel_l2_outside(tiled, whole, band) < 5e-3, rel_l2_outside(tiled, whole, band)
    assert st["pre_max"] == pytest.approx(st1["pre_max"], rel=2e-3) and st["norm_function"] == st1["norm_function"]
    if (world, h, w) == (2, 32, 8):
        assert _rel(tiled, whole) < 1e-6, _rel(tiled, whole)


def test_full_size_c2_properties_and_gpu_oracle(setup):
    """BASELINE config C2 at full size (4 x 16x128x128 -> 4 x 1024^2, "moderate"): properties that do not need a CPU
    oracle run, plus the fp32 PyTorch oracle executed on the same GPU (TF32 off) for the whole batch.
      * the multiplier is applied last (hdr_vae_decode.py:180-182): decode(ev=2) == 2 * decode(ev=1) exactly;
      * batch order: decoding a permuted batch permutes the images (statistics are batch-global sums / extrema);
        identical up to the order of the fp64 partial sums;
      * rel-L2 <= 1e-2 against the fp32 oracle on all 4 images, statistics in agreement."""
    dec, eng = setup
    z = make_latent(4, 128, 128, seed=1234).to(DEV)
    out1, st1 = eng.decode(z, "moderate", 1.0)
    out2, _ = eng.decode(z, "moderate", 2.0)
    assert out1.shape == (4, 1024, 1024, 3)
    assert torch.equal(out2, out1 * 2.0)
    perm = torch.tensor([2, 0, 3, 1], device=DEV)
    outp, _ = eng.decode(z[perm].contiguous(), "moderate", 1.0)
    assert torch.allclose(outp, out1[perm], rtol=1e-6, atol=1e-7)
    del out2, outp
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dec_gpu = dec.to(DEV)
    try:
        ref, rst, _ = ho.simple_hdr_decode(dec_gpu, z, "moderate", 1.0)
    finally:
        dec.to("cpu")
    assert _rel(out1, ref) < 1e-2, _rel(out1, ref)
    for b in range(4):
        assert _rel(out1[b], ref[b]) < 1e-2
    assert st1["hdr_pixels"] == pytest.approx(rst["hdr_pixels"], rel=2e-3)
    assert st1["out_max"] == pytest.approx(rst["out_max"], rel=2e-2)


def test_large_2048_decode_vs_gpu_oracle_and_row_tiling(setup):
    """Towards config C4 (the largest size whose fp32 oracle still runs in seconds on the GPU): 1x16x256x256 latent ->
    2048^2, "aggressive" (= mathematical_recovery).  T = 65 536 tokens exercises the chunked / split-K attention.
      * rel-L2 <= 1e-2 against the fp32 PyTorch oracle on the same GPU (TF32 off);
      * the same image row-tiled over 4 virtual ranks (64 latent rows each: conv tiles coincide) is identical."""
    from vae_decode_hdr_b200.sharding import decode_rows_emulated
    dec, eng = setup
    z = make_latent(1, 256, 256, seed=1234).to(DEV)
    out, st = eng.decode(z, "aggressive", 1.0)
    assert out.shape == (1, 2048, 2048, 3) and bool(torch.isfinite(out).all())
    tiled, _ = decode_rows_emulated(eng, z, 4, "aggressive")
    assert _rel(tiled, out) < 1e-6, _rel(tiled, out)
    del tiled
    eng._workspace = None
    torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dec_gpu = dec.to(DEV)
    try:
        ref, rst, _ = ho.simple_hdr_decode(dec_gpu, z, "aggressive", 1.0)
    finally:
        dec.to("cpu")
    assert _rel(out, ref) < 1e-2, _rel(out, ref)
    assert st["pre_max"] == pytest.approx(rst["pre_max"], rel=5e-3)


_SWITCH_CHILD = r"""
import sys, torch
sys.path.insert(0, ".")
from oracle.flux_decoder import build_decoder, make_latent
from vae_decode_hdr_b200.engine import HdrVaeEngine
eng = HdrVaeEngine(build_decoder(0).state_dict(), "cuda:0")
z = make_latent(2, 16, 24, seed=5).to("cuda:0")
img, _ = eng.decode(z, "moderate")
feat = eng.decode_features(z).float()
torch.save({"img": img.cpu(), "feat": feat.cpu()}, sys.argv[1])
"""


def _decode_in_child(tmp_path, name, env_extra):
    """The numeric-plan switches are read once per process, so the other setting runs in a child process."""
    import os
    import subprocess
    import sys
    out = tmp_path / name
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", _SWITCH_CHILD, str(out)], env=env, capture_output=True, text=True, timeout=300,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(out)


@pytest.mark.gpu
def test_groupnorm_inside_the_conv_is_bit_identical_to_the_streaming_kernel(tmp_path):
    """The transform warps of the 256-column convs (gemm_tc.cu, XF builds) perform gn_apply_kernel's operations in its
    order on the slab in shared memory: features and image of a decode must not change by one bit when HDRVAE_FUSE_GN=0
    sends every GroupNorm through the streaming kernel instead."""
    fused = _decode_in_child(tmp_path, "fused.pt", {})
    streamed = _decode_in_child(tmp_path, "streamed.pt", {"HDRVAE_FUSE_GN": "0"})
    assert torch.equal(fused["feat"], streamed["feat"])
    assert torch.equal(fused["img"], streamed["img"])


@pytest.mark.gpu
def test_fp16_residual_stream_against_the_fp32_stream_and_the_oracle(tmp_path):
    """Default numeric plan (residual stream stored as scaled fp16) vs HDRVAE_X16=0 (fp32 stream) vs the fp32 oracle: both
    within the 1e-2 image bound, and the cost of the 16-bit stream on the features stays below 1e-3 rel-L2."""
    dec = build_decoder(0)
    z = make_latent(2, 16, 24, seed=5)
    with torch.no_grad():
        ref, _, _ = ho.simple_hdr_decode(dec, z, "moderate", 1.0)
        fref = dec.features(z).permute(0, 2, 3, 1)
    x16 = _decode_in_child(tmp_path, "x16.pt", {})
    x32 = _decode_in_child(tmp_path, "x32.pt", {"HDRVAE_X16": "0"})
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    e16, e32 = rel(x16["feat"], fref), rel(x32["feat"], fref)
    assert e32 < 3e-3 and e16 < 4e-3 and e16 - e32 < 1e-3, (e16, e32)
    assert rel(x16["img"], ref) < 1e-2 and rel(x32["img"], ref) < 1e-2
