"""CPU tests of the upscaler oracle (oracle/upscaler_oracle.py): golden vectors produced by the UNMODIFIED reference
node (oracle/make_golden_upscale.py), the restated third-party pieces against independent formulations, and — when
/root/reference is mounted — the live reference node on a fresh seed."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import upscaler_oracle as uo
from oracle.make_golden_upscale import CASES, fingerprint, make_image
from oracle.ref_loader import reference_available


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))


@pytest.mark.parametrize("case", sorted(CASES))
def test_upscale_oracle_matches_reference_golden(golden_dir, case):
    g = _load(golden_dir, case)
    b, h, w, seed, gain, nb, last_gain, blur, fix, method = CASES[case]
    net = uo.build_upscaler(0, nb=nb, gain=last_gain)
    assert fingerprint(net) == pytest.approx(float(g["weight_fingerprint"]), rel=1e-12)
    img = make_image(b, h, w, seed, gain)
    assert np.array_equal(img.numpy(), g["image"])
    out = uo.upscale(img, uo.FakeDescriptor(net), blur, fix, method)
    ref = torch.from_numpy(g["output"])
    # same machine, same torch: bit-exact; across BLAS / thread counts the convs may reorder sums
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_tile_positions_and_steps():
    assert uo.tile_positions(512, 512, 64) == [0]
    assert uo.tile_positions(520, 512, 64) == [0, 448]
    assert uo.tile_positions(1024, 512, 64) == [0, 448, 896]
    assert uo.get_tiled_scale_steps(1024, 1024, 512, 512, 64) == 9
    assert uo.get_tiled_scale_steps(512, 300, 512, 512, 64) == 1


def test_tiled_scale_of_pointwise_function_is_exact_resampling():
    """For a function without spatial context the feather-blended result equals the untiled one."""
    x = torch.rand(1, 3, 70, 45)
    f = lambda a: F.interpolate(a, scale_factor=2, mode="nearest") * 0.5 + 0.25   # noqa: E731
    tiled = uo.tiled_scale(x, f, tile_x=32, tile_y=32, overlap=8, upscale_amount=2)
    assert torch.allclose(tiled, f(x), atol=1e-6)


def test_median_blur_matches_sort_based_median():
    x = torch.randn(2, 3, 9, 7)
    ref = torch.zeros_like(x)
    xp = F.pad(x, (1, 1, 1, 1))
    for i in range(9):
        for j in range(7):
            ref[:, :, i, j] = xp[:, :, i:i + 3, j:j + 3].reshape(2, 3, 9).sort(dim=-1).values[..., 4]
    assert torch.equal(uo.median_blur(x, (3, 3)), ref)


def test_ycbcr_roundtrip_is_close_to_identity():
    x = torch.rand(1, 3, 8, 8) * 4 - 1
    assert torch.allclose(uo.ycbcr_to_rgb(uo.rgb_to_ycbcr(x)), x, atol=5e-3)


def test_old_arch_key_mapping_covers_every_tensor():
    net = uo.build_upscaler(0, nb=2)
    sd = uo.state_dict_old_arch(net)
    assert len(sd) == len(net.state_dict())
    assert "model.0.weight" in sd and "model.1.sub.2.weight" in sd and "model.1.sub.1.RDB3.conv5.0.bias" in sd
    assert "model.10.bias" in sd


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_live_reference_upscaler_agrees_with_oracle():
    from oracle.ref_loader import load_reference_upscaler
    net = uo.build_upscaler(3, nb=1, gain=25.0)
    desc = uo.FakeDescriptor(net)
    img = make_image(1, 18, 530, 21, 3.0)
    node = load_reference_upscaler(desc)
    with torch.no_grad():
        (ref,) = node.upscale(img, "fake_esrgan.pth", False, True, "nearest-exact")
    assert torch.equal(ref, uo.upscale(img, desc, False, True, "nearest-exact"))
