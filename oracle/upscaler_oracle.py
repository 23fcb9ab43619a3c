"""CPU restatement of the reference's HDR upscaler path (`HDRUpscaleWithModel.upscale`,
/root/reference/hdr_upscale_with_model.py:148-263) and of the third-party pieces it calls.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU baseline leg, never by
the product path.

What is pinned and what is not
------------------------------
* The node's own glue (hook, two passes, YCbCr recombination, median, local fix; hdr_upscale_with_model.py) is
  pinned: `oracle/make_golden_upscale.py` imports the UNMODIFIED reference file with stand-ins for the absent
  packages and checks `upscale()` below against it bit for bit (tests/test_oracle_cpu.py repeats that whenever
  /root/reference is present) and stores golden vectors.
* The arithmetic that lives in packages that are NOT under /root/reference and not installed here is restated
  from their published algorithms — **parity unpinned** for these:
    - spandrel (unlisted, unpinned; hdr_upscale_with_model.py:6): ESRGAN / RRDBNet architecture
      (nf 64, nb 23, gc 32, scale 4, LeakyReLU 0.2, residual scaling 0.2, nearest 2x before each up-conv);
    - ComfyUI `comfy.utils.tiled_scale` / `get_tiled_scale_steps` / `common_upscale` (unpinned): overlapping
      tiles, linear feather of `overlap * scale` pixels on every tile edge, out / out_div;
    - kornia >= 0.8.1 `rgb_to_ycbcr` (Y = .299R + .587G + .114B, Cb = (B - Y) * .564 + .5,
      Cr = (R - Y) * .713 + .5) and `median_blur` (3x3, zero padded).
  torchvision (`gaussian_blur`, hdr_upscale_with_model.py:7) IS installed and is called directly.
"""
from __future__ import annotations

import itertools
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- RRDBNet (ESRGAN, as spandrel builds it)
class _RDB(nn.Module):
    def __init__(self, nf=64, gc=32):
        super().__init__()
        self.conv1 = nn.Conv2d(nf, gc, 3, 1, 1)
        self.conv2 = nn.Conv2d(nf + gc, gc, 3, 1, 1)
        self.conv3 = nn.Conv2d(nf + 2 * gc, gc, 3, 1, 1)
        self.conv4 = nn.Conv2d(nf + 3 * gc, gc, 3, 1, 1)
        self.conv5 = nn.Conv2d(nf + 4 * gc, nf, 3, 1, 1)

    def forward(self, x):
        x1 = F.leaky_relu(self.conv1(x), 0.2)
        x2 = F.leaky_relu(self.conv2(torch.cat((x, x1), 1)), 0.2)
        x3 = F.leaky_relu(self.conv3(torch.cat((x, x1, x2), 1)), 0.2)
        x4 = F.leaky_relu(self.conv4(torch.cat((x, x1, x2, x3), 1)), 0.2)
        x5 = self.conv5(torch.cat((x, x1, x2, x3, x4), 1))
        return x5 * 0.2 + x


class _RRDB(nn.Module):
    def __init__(self, nf=64, gc=32):
        super().__init__()
        self.rdb1, self.rdb2, self.rdb3 = _RDB(nf, gc), _RDB(nf, gc), _RDB(nf, gc)

    def forward(self, x):
        return self.rdb3(self.rdb2(self.rdb1(x))) * 0.2 + x


class RRDBNet(nn.Module):
    """ESRGAN generator, 4x.  Parameter names follow the Real-ESRGAN ("new arch") checkpoint layout; see
    `old_arch_key` for the BasicSR / spandrel ("old arch") names of the same tensors."""

    def __init__(self, nf=64, nb=23, gc=32, in_ch=3, out_ch=3):
        super().__init__()
        self.nf, self.nb, self.gc = nf, nb, gc
        self.conv_first = nn.Conv2d(in_ch, nf, 3, 1, 1)
        self.body = nn.Sequential(*[_RRDB(nf, gc) for _ in range(nb)])
        self.conv_body = nn.Conv2d(nf, nf, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.conv_hr = nn.Conv2d(nf, nf, 3, 1, 1)
        self.conv_last = nn.Conv2d(nf, out_ch, 3, 1, 1)

    def forward(self, x):
        feat = self.conv_first(x)
        feat = feat + self.conv_body(self.body(feat))
        feat = F.leaky_relu(self.conv_up1(F.interpolate(feat, scale_factor=2, mode="nearest")), 0.2)
        feat = F.leaky_relu(self.conv_up2(F.interpolate(feat, scale_factor=2, mode="nearest")), 0.2)
        return self.conv_last(F.leaky_relu(self.conv_hr(feat), 0.2))


def old_arch_key(key: str, nb: int = 23) -> str:
    """Real-ESRGAN name -> the `model.N...` name spandrel's ESRGAN module uses for the same tensor."""
    head, _, leaf = key.rpartition(".")
    table = {"conv_first": "model.0", "conv_body": f"model.1.sub.{nb}", "conv_up1": "model.3", "conv_up2": "model.6",
             "conv_hr": "model.8", "conv_last": "model.10"}
    if head in table:
        return f"{table[head]}.{leaf}"
    _, i, rdb, conv = head.split(".")                      # body.i.rdbK.convJ
    return f"model.1.sub.{i}.RDB{rdb[-1]}.{conv}.0.{leaf}"


def build_upscaler(seed: int = 0, nb: int = 23, gain: float = 1.0) -> RRDBNet:
    """Random-init RRDBNet (PyTorch default init under `seed`, BASELINE config C5).  `gain` scales conv_last so
    that tests can push the output towards the +-1 saturation of the atanh hook."""
    torch.manual_seed(seed)
    net = RRDBNet(nb=nb).eval()
    if gain != 1.0:
        with torch.no_grad():
            net.conv_last.weight.mul_(gain)
            net.conv_last.bias.mul_(gain)
    for p in net.parameters():
        p.requires_grad_(False)
    return net


class _Arch:
    def __init__(self, name):
        self.name = name


class FakeDescriptor:
    """Stand-in for spandrel's ImageModelDescriptor: `.model`, `.scale`, `.architecture.name`."""

    def __init__(self, model: nn.Module, scale: int = 4, arch: str = "ESRGAN"):
        self.model, self.scale, self.architecture = model, scale, _Arch(arch)


# ----------------------------------------------------------------------------- ComfyUI comfy.utils restated
def get_tiled_scale_steps(width, height, tile_x, tile_y, overlap):
    rows = 1 if height <= tile_y else -(-(height - overlap) // (tile_y - overlap))
    cols = 1 if width <= tile_x else -(-(width - overlap) // (tile_x - overlap))
    return rows * cols


def tile_positions(size: int, tile: int, overlap: int):
    """Start offsets of the tiles along one axis (comfy.utils.tiled_scale_multidim)."""
    if size <= tile:
        return [0]
    return list(range(0, size - overlap, tile - overlap))


def tiled_scale(samples, function, tile_x=64, tile_y=64, overlap=8, upscale_amount=4, out_channels=3, pbar=None):
    """comfy.utils.tiled_scale: run `function` on overlapping tiles, blend with a linear feather mask."""
    tile = (tile_y, tile_x)
    scale = upscale_amount
    B, _, H, W = samples.shape
    output = torch.empty((B, out_channels, round(H * scale), round(W * scale)), dtype=samples.dtype)
    for b in range(B):
        s = samples[b:b + 1]
        if H <= tile[0] and W <= tile[1]:                   # the whole image fits one tile: no mask at all
            output[b:b + 1] = function(s)
            continue
        out = torch.zeros((1, out_channels, round(H * scale), round(W * scale)), dtype=samples.dtype)
        out_div = torch.zeros_like(out)
        pos_lists = [tile_positions(s.shape[d + 2], tile[d], overlap) for d in range(2)]
        for it in itertools.product(*pos_lists):
            s_in = s
            upscaled = []
            for d in range(2):
                pos = max(0, min(s.shape[d + 2] - overlap, it[d]))
                length = min(tile[d], s.shape[d + 2] - pos)
                s_in = s_in.narrow(d + 2, pos, length)
                upscaled.append(round(pos * scale))
            ps = function(s_in)
            mask = torch.ones_like(ps)
            feather = round(overlap * scale)
            for d in range(2, 4):
                if feather >= mask.shape[d]:
                    continue
                for t in range(feather):
                    a = (t + 1) / feather
                    mask.narrow(d, t, 1).mul_(a)
                    mask.narrow(d, mask.shape[d] - 1 - t, 1).mul_(a)
            o = out.narrow(2, upscaled[0], mask.shape[2]).narrow(3, upscaled[1], mask.shape[3])
            o_d = out_div.narrow(2, upscaled[0], mask.shape[2]).narrow(3, upscaled[1], mask.shape[3])
            o.add_(ps * mask)
            o_d.add_(mask)
        output[b:b + 1] = out / out_div
    return output


def bislerp(samples, width, height):
    """ComfyUI's own "bislerp" resampler (comfy.utils.bislerp; third party, un-vendored, un-pinned — restated from the
    published source, parity unpinned): separable (width pass, then height pass) SPHERICAL linear interpolation of
    the per-pixel channel vectors between the two source pixels that bilinear interpolation (align_corners=False)
    would blend, with the blend ratio of that bilinear interpolation.  For the single-channel luma the reference
    feeds it (hdr_upscale_with_model.py:235-240) the vectors are scalars: same sign -> the FIRST pixel is returned
    (dot = 1 > 1 - 1e-5), opposite signs -> plain linear interpolation, a zero -> the sine-weighted formula."""
    def slerp(b1, b2, r):
        c = b1.shape[-1]
        b1_norms = torch.norm(b1, dim=-1, keepdim=True)
        b2_norms = torch.norm(b2, dim=-1, keepdim=True)
        b1_normalized = b1 / b1_norms
        b2_normalized = b2 / b2_norms
        b1_normalized[b1_norms.expand(-1, c) == 0.0] = 0.0
        b2_normalized[b2_norms.expand(-1, c) == 0.0] = 0.0
        dot = (b1_normalized * b2_normalized).sum(1)
        omega = torch.acos(dot)
        so = torch.sin(omega)
        res = (torch.sin((1.0 - r.squeeze(1)) * omega) / so).unsqueeze(1) * b1_normalized + \
              (torch.sin(r.squeeze(1) * omega) / so).unsqueeze(1) * b2_normalized
        res *= (b1_norms * (1.0 - r) + b2_norms * r).expand(-1, c)
        res[dot > 1 - 1e-5] = b1[dot > 1 - 1e-5]
        res[dot < 1e-5 - 1] = (b1 * (1.0 - r) + b2 * r)[dot < 1e-5 - 1]
        return res

    def generate_bilinear_data(length_old, length_new, device):
        coords_1 = torch.arange(length_old, dtype=torch.float32, device=device).reshape((1, 1, 1, -1))
        coords_1 = F.interpolate(coords_1, size=(1, length_new), mode="bilinear")
        ratios = coords_1 - coords_1.floor()
        coords_1 = coords_1.to(torch.int64)
        coords_2 = torch.arange(length_old, dtype=torch.float32, device=device).reshape((1, 1, 1, -1)) + 1
        coords_2[:, :, :, -1] -= 1
        coords_2 = F.interpolate(coords_2, size=(1, length_new), mode="bilinear")
        coords_2 = coords_2.to(torch.int64)
        return ratios, coords_1, coords_2

    orig_dtype = samples.dtype
    samples = samples.float()
    n, c, h, w = samples.shape
    h_new, w_new = (height, width)

    ratios, coords_1, coords_2 = generate_bilinear_data(w, w_new, samples.device)
    coords_1 = coords_1.expand((n, c, h, -1))
    coords_2 = coords_2.expand((n, c, h, -1))
    ratios = ratios.expand((n, 1, h, -1))
    pass_1 = samples.gather(-1, coords_1).movedim(1, -1).reshape((-1, c))
    pass_2 = samples.gather(-1, coords_2).movedim(1, -1).reshape((-1, c))
    ratios = ratios.movedim(1, -1).reshape((-1, 1))
    result = slerp(pass_1, pass_2, ratios)
    result = result.reshape(n, h, w_new, c).movedim(-1, 1)

    ratios, coords_1, coords_2 = generate_bilinear_data(h, h_new, samples.device)
    coords_1 = coords_1.reshape((1, 1, -1, 1)).expand((n, c, -1, w_new))
    coords_2 = coords_2.reshape((1, 1, -1, 1)).expand((n, c, -1, w_new))
    ratios = ratios.reshape((1, 1, -1, 1)).expand((n, 1, -1, w_new))
    pass_1 = result.gather(-2, coords_1).movedim(1, -1).reshape((-1, c))
    pass_2 = result.gather(-2, coords_2).movedim(1, -1).reshape((-1, c))
    ratios = ratios.movedim(1, -1).reshape((-1, 1))
    result = slerp(pass_1, pass_2, ratios)
    result = result.reshape(n, h_new, w_new, c).movedim(-1, 1)
    return result.to(orig_dtype)


def common_upscale(samples, width, height, upscale_method, crop):
    """comfy.utils.common_upscale (crop "disabled"/False): torch's interpolate for the methods torch offers, ComfyUI's
    own bislerp otherwise."""
    if upscale_method == "bislerp":
        return bislerp(samples, width, height)
    return F.interpolate(samples, size=(height, width), mode=upscale_method)


# ----------------------------------------------------------------------------- kornia restated
def rgb_to_ycbcr(image):
    r, g, b = image[..., 0, :, :], image[..., 1, :, :], image[..., 2, :, :]
    delta = 0.5
    y = 0.299 * r + 0.587 * g + 0.114 * b
    cb = (b - y) * 0.564 + delta
    cr = (r - y) * 0.713 + delta
    return torch.stack([y, cb, cr], -3)


def median_blur(x, kernel_size=(3, 3)):
    """kornia.filters.median_blur: zero-padded window, median over the kh*kw samples."""
    kh, kw = kernel_size
    b, c, h, w = x.shape
    patches = F.unfold(x.reshape(b * c, 1, h, w), (kh, kw), padding=(kh // 2, kw // 2))     # zero padding
    return patches.median(dim=1).values.reshape(b, c, h, w)


# ----------------------------------------------------------------------------- the node's own glue, restated
def ycbcr_to_rgb(image):
    """hdr_upscale_with_model.py:20-48 (the reference's un-clamped variant)."""
    y, cb, cr = image[..., 0, :, :], image[..., 1, :, :], image[..., 2, :, :]
    cb_s, cr_s = cb - 0.5, cr - 0.5
    r = y + 1.403 * cr_s
    g = y - 0.714 * cr_s - 0.344 * cb_s
    b = y + 1.773 * cb_s
    return torch.stack([r, g, b], -3)


def reversal_kind(arch_name: str) -> str:
    """hdr_upscale_with_model.py:266-279."""
    if arch_name in ("ESRGAN", "RealESRGAN", "SwinIR", "HAT") or "VAE" in arch_name:
        return "atanh"
    return "logit"


def reversal(output, kind: str):
    """hdr_upscale_with_model.py:79-107 (forward hook on the model: applied to every tile's output)."""
    if kind == "logit":
        return torch.logit(torch.clamp(output, 1e-7, 1 - 1e-7))
    return torch.atanh(torch.clamp(output, -1 + 1e-6, 1 - 1e-6))


def upscale(image_bhwc, descriptor, small_blur=False, local_fix=False, upscale_method="bilinear", tile=512, overlap=64):
    """hdr_upscale_with_model.py:148-263 on CPU.  image: [B,H,W,3] fp32 -> [B, scale*H, scale*W, 3] fp32."""
    model, scale = descriptor.model, descriptor.scale
    kind = reversal_kind(descriptor.architecture.name)
    x = image_bhwc.movedim(-1, -3)
    if small_blur:
        from torchvision.transforms.functional import gaussian_blur
        x = gaussian_blur(x, kernel_size=3, sigma=0.1)

    def run(inp):
        with torch.no_grad():
            return tiled_scale(inp, lambda a: reversal(model(a), kind), tile_x=tile, tile_y=tile, overlap=overlap,
                               upscale_amount=scale)
    s_unclamped = run(x)
    s_clamped = run(torch.clamp(x, -1.0, 1.0))
    yc_c, yc_u = rgb_to_ycbcr(s_clamped), rgb_to_ycbcr(s_unclamped)
    y = median_blur(torch.clamp(yc_u[:, 0:1], min=0.0, max=8.0), (3, 3))
    s_final = ycbcr_to_rgb(torch.cat([y, yc_c[:, 1:2], yc_c[:, 2:3]], dim=1))
    if small_blur:
        s_final = median_blur(s_final, (3, 3))
    if local_fix:
        y_orig = rgb_to_ycbcr(x)[:, 0:1]
        y_scaled = common_upscale(y_orig, s_final.shape[3], s_final.shape[2], upscale_method, False)
        mask = (y_scaled < 0.1).float()
        s_final = s_final * (1.0 - mask) + torch.clamp(s_final, -1.0, 1.0) * mask
    return s_final.movedim(-3, -1)


def state_dict_old_arch(net: RRDBNet) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((old_arch_key(k, net.nb), v) for k, v in net.state_dict().items())
