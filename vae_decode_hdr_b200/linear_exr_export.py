"""`LinearEXRExport` — drop-in for the reference node of the same name
(/root/reference/linear_exr_export.py:81-118 interface, :230-369 behaviour): same INPUT_TYPES / RETURN_TYPES /
FUNCTION / CATEGORY / OUTPUT_NODE, same filename rules (version `_vNNN` by directory scan :43-78,:290-295, frame
`_frame_%0Nd` :297-299, leading "/" = sub-folder of the ComfyUI output directory :268-273), same sidecar workflow JSON
(:119-143), same error convention (every failure is returned as the string "ERROR: ...", :366-369).

What changes is the data path.  The reference casts on the CPU with `ndarray.astype(np.float16)` (:155,165) and hands
the array to pyexr / imageio / cv2 — none of which is a dependency here.  This node packs on the GPU
(`hdrvae_pack_half`, bit-exact with the numpy cast: round-to-nearest-even, overflow -> inf) directly in OpenEXR
scan-line order (per image row the B, G, R planes), so the packed buffer IS the pixel payload of the file, and writes
the container itself: single-part scan-line OpenEXR 2.0 with HALF or FLOAT channels, compression `none` or `zip`
(`rle` / `piz` / `pxr24` requests are stored as `zip`: the pixels are identical, only the storage differs; pxr24 would
be lossy for 32-bit data), or flat Radiance RGBE for `format="hdr"`.  Container layout restated from the published
OpenEXR file-format specification ("OpenEXR File Layout") and the Radiance picture format; no third-party code.
"""
from __future__ import annotations

import json
import logging
import os
import re
import struct
import traceback
import zlib
from glob import glob
from typing import Dict, Optional, Tuple

import numpy as np
import torch

logger = logging.getLogger("LinearEXRExport")

_EXR_COMPRESSION = {"none": 0, "zip": 3}
_PIXEL_HALF, _PIXEL_FLOAT = 1, 2


def highest_version(directory: str, prefix: str) -> int:
    """Largest N among files `<prefix>_vN...` in `directory` (0 if none) — linear_exr_export.py:43-78."""
    rx = re.compile(r"^" + re.escape(prefix) + r"_v(\d+).*$")
    best = 0
    for path in glob(os.path.join(directory, f"{prefix}*")):
        m = rx.match(os.path.basename(path))
        if m:
            best = max(best, int(m.group(1)))
    return best


# ---------------------------------------------------------------------------------------------- containers
def _attr(name: str, typ: str, payload: bytes) -> bytes:
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def _zip_block(raw: np.ndarray) -> bytes:
    """OpenEXR ZIP block: bytes de-interleaved (even positions, then odd), delta-predicted (+128), deflated; stored
    raw when deflate does not shrink it."""
    n = raw.size
    tmp = np.empty(n, dtype=np.uint8)
    half = (n + 1) // 2
    tmp[:half] = raw[0::2]
    tmp[half:] = raw[1::2]
    pred = tmp.copy()
    pred[1:] = (tmp[1:].astype(np.int16) - tmp[:-1].astype(np.int16) + 128).astype(np.uint8)
    comp = zlib.compress(pred.tobytes(), 6)
    return comp if len(comp) < n else raw.tobytes()


def write_exr_scanlines(path: str, planes: np.ndarray, compression: str = "zip") -> None:
    """planes: [H, 3, W] float16 or float32, the B, G, R planes of every row (= OpenEXR's channel order inside a
    scan line: alphabetical).  Writes a single-part scan-line OpenEXR file."""
    if planes.ndim != 3 or planes.shape[1] != 3 or planes.dtype not in (np.float16, np.float32):
        raise ValueError(f"expected [H,3,W] float16/float32 planes, got {planes.shape} {planes.dtype}")
    H, _, W = planes.shape
    comp = _EXR_COMPRESSION["none" if compression == "none" else "zip"]
    ptype = _PIXEL_HALF if planes.dtype == np.float16 else _PIXEL_FLOAT
    chlist = b"".join(c + b"\0" + struct.pack("<iB3xii", ptype, 0, 1, 1) for c in (b"B", b"G", b"R")) + b"\0"
    box = struct.pack("<iiii", 0, 0, W - 1, H - 1)
    header = (b"\x76\x2f\x31\x01" + struct.pack("<i", 2) +
              _attr("channels", "chlist", chlist) + _attr("compression", "compression", bytes([comp])) +
              _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box) +
              _attr("lineOrder", "lineOrder", b"\0") + _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) +
              _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)) +
              _attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    lines = 16 if comp == 3 else 1
    raw = np.ascontiguousarray(planes).view(np.uint8).reshape(H, -1)
    blocks = []
    for y0 in range(0, H, lines):
        blk = raw[y0:y0 + lines].reshape(-1)
        blocks.append((y0, _zip_block(blk) if comp == 3 else blk.tobytes()))
    offset = len(header) + 8 * len(blocks)
    table = []
    for _, data in blocks:
        table.append(offset)
        offset += 8 + len(data)
    with open(path, "wb") as f:
        f.write(header)
        f.write(struct.pack(f"<{len(table)}Q", *table))
        for y0, data in blocks:
            f.write(struct.pack("<ii", y0, len(data)))
            f.write(data)


def write_radiance_hdr(path: str, rgb: np.ndarray) -> None:
    """[H,W,3] float32 -> flat (un-run-length-encoded) Radiance RGBE picture; negative values clamp to 0 as RGBE has
    no sign."""
    H, W, _ = rgb.shape
    v = np.maximum(rgb.astype(np.float32), 0.0)
    m = v.max(axis=-1)
    mant, expo = np.frexp(m)
    scale = np.where(m > 1e-32, mant * 256.0 / np.maximum(m, 1e-38), 0.0).astype(np.float32)
    out = np.zeros((H, W, 4), dtype=np.uint8)
    out[..., :3] = np.clip(v * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, expo + 128, 0).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {H} +X {W}\n".encode())
        f.write(out.tobytes())


# ---------------------------------------------------------------------------------------------- the node
class LinearEXRExport:
    """Linear EXR Export (B200-native pack).  Drop-in for the reference node class of the same name."""
    VERSION_TRACKER: Dict[str, int] = {}

    @classmethod
    def INPUT_TYPES(cls):
        # linear_exr_export.py:90-112
        return {
            "required": {
                "hdr_image": ("IMAGE", {"tooltip": "HDR image tensor with values potentially above 1.0"}),
                "filename_prefix": ("STRING", {"default": "comfyUI", "tooltip": "Base filename (without extension)"}),
            },
            "optional": {
                "versioning": ("BOOLEAN", {"default": False, "tooltip": "Incremental versioning save. adding v001, v002... to it's file name"}),
                "frame_sequence": ("BOOLEAN", {"default": False, "tooltip": "Save animation into multiple frames 1001, 1002..."}),
                "start_frame": ("INT", {"default": 1001, "min": 0, "max": 99999999}),
                "frame_pad": ("INT", {"default": 4, "min": 1, "max": 8}),
                "output_path": ("STRING", {"default": "/HDR", "tooltip": "Output path: Empty=default ComfyUI/output, /subfolder=output/subfolder, or full custom path"}),
                "format": (["exr", "hdr"], {"default": "exr", "tooltip": "file format"}),
                "bit_depth": (["16bit", "32bit"], {"default": "16bit", "tooltip": "EXR precision: 32bit = maximum quality, 16bit = smaller files"}),
                "compression": (["none", "rle", "zip", "piz", "pxr24"], {"default": "zip", "tooltip": "EXR compression type"}),
                "save_workflow": ("BOOLEAN", {"default": False, "tooltip": "Saves the workflow JSON to a sidecar file next to the HDR image"}),
            },
            "hidden": {
                "prompt": "PROMPT",
                "extra_pnginfo": "EXTRA_PNGINFO",
            },
        }

    RETURN_TYPES = ("STRING",)
    RETURN_NAMES = ("filepath",)
    FUNCTION = "export_linear_exr"
    CATEGORY = "image"
    OUTPUT_NODE = True

    # -- helpers ---------------------------------------------------------------------------------------
    @staticmethod
    def _output_directory() -> str:
        """ComfyUI's output directory (folder_paths), else `<ComfyUI root>/output` found by walking up from this
        package, else two levels up — linear_exr_export.py:431-472."""
        try:
            import folder_paths
            return folder_paths.get_output_directory()
        except Exception:
            here = os.path.dirname(os.path.abspath(__file__))
            d = here
            for _ in range(5):
                if os.path.exists(os.path.join(d, "custom_nodes")) and os.path.exists(os.path.join(d, "models")):
                    return os.path.join(d, "output")
                d = os.path.dirname(d)
            return os.path.join(os.path.dirname(os.path.dirname(here)), "output")

    @staticmethod
    def _sidecar(filepath: str, prompt: Optional[dict], extra_pnginfo: Optional[dict]) -> None:
        doc = {"prompt": prompt or {}, "extra_pnginfo": extra_pnginfo or {}}
        if doc["prompt"] or doc["extra_pnginfo"]:
            with open(os.path.splitext(filepath)[0] + ".json", "w") as f:
                json.dump(doc, f, indent=4)

    @staticmethod
    def _pack_frames(hdr_image: torch.Tensor, half: bool) -> np.ndarray:
        """[B,H,W,3] -> host [B,H,3,W] B/G/R planes per row, float16 (GPU pack kernel) or float32."""
        from .engine import pack_half
        if half:
            if not torch.cuda.is_available():
                raise RuntimeError("LinearEXRExport (B200): the half pack runs on the GPU (hdrvae_pack_half); no CUDA device")
            img = hdr_image if hdr_image.device.type == "cuda" else hdr_image.to("cuda", non_blocking=True)
            packed = pack_half(img.float(), exr_scanline_order=True)
            host = torch.empty(packed.shape, dtype=packed.dtype, pin_memory=True)
            host.copy_(packed, non_blocking=True)
            torch.cuda.current_stream(packed.device).synchronize()
            return host.numpy()
        arr = hdr_image.detach().float().cpu().numpy()
        return np.ascontiguousarray(arr[..., ::-1].transpose(0, 1, 3, 2))

    def export_linear_exr(self, hdr_image: torch.Tensor, filename_prefix: str = "HDR_VAE", output_path: str = "",
                          start_frame: int = 1, frame_pad: int = 4, versioning: bool = True, frame_sequence: bool = False,
                          format: str = "hdr", bit_depth: str = "16bit", compression: str = "zip",
                          save_workflow: bool = False, prompt: dict = None, extra_pnginfo: dict = None) -> Tuple[str]:
        try:
            if hdr_image.dim() == 3:
                hdr_image = hdr_image.unsqueeze(0)
            if hdr_image.dim() != 4 or hdr_image.shape[-1] != 3:
                raise ValueError(f"expected an IMAGE tensor [B,H,W,3], got {tuple(hdr_image.shape)}")
            fmt = str(format).lower()
            if fmt not in ("exr", "hdr"):
                raise ValueError(f"Unsupported format: {format}")
            batch = hdr_image.shape[0]

            clean = output_path.strip() if output_path else ""
            if not clean:
                out_dir = self._output_directory()
            elif clean.startswith("/"):
                out_dir = os.path.join(self._output_directory(), clean[1:])
            else:
                out_dir = clean
            parts = filename_prefix.replace("/", os.sep).replace("\\", os.sep).split(os.sep)
            base = parts[-1]
            if len(parts) > 1:
                out_dir = os.path.join(out_dir, *parts[:-1])
            os.makedirs(out_dir, exist_ok=True)

            name = base
            if versioning:
                name += f"_v{highest_version(os.path.normpath(out_dir), base) + 1:03d}"
            numbered = batch > 1 or frame_sequence
            if numbered:
                name += f"_frame_%0{frame_pad}d"
            name += f".{fmt}"

            frames = None
            if fmt == "exr":
                frames = self._pack_frames(hdr_image, half=(bit_depth != "32bit"))
            else:
                rgb = hdr_image.detach().float().cpu().numpy()
            last = None
            for i in range(batch):
                path = os.path.join(out_dir, name % (start_frame + i) if numbered else name)
                if fmt == "exr":
                    write_exr_scanlines(path, frames[i], compression)
                else:
                    write_radiance_hdr(path, rgb[i])
                if i == 0 and save_workflow:
                    self._sidecar(path, prompt, extra_pnginfo)
                last = path
            if last is None:
                raise RuntimeError("Export completed, but no file paths were recorded.")
            logger.info("Linear %s exported: %d frames, last %s (%.2f MB)", fmt.upper(), batch, last,
                        os.path.getsize(last) / 2**20)
            return (last,)
        except Exception as e:                       # the reference's convention: report, do not raise (:366-369)
            logger.error("Linear EXR export failed: %s\n%s", e, traceback.format_exc())
            return (f"ERROR: {str(e)}",)
