"""`HDRUpscaleWithModel` — drop-in for the reference node of the same name
(/root/reference/hdr_upscale_with_model.py:50-279): same INPUT_TYPES / RETURN_TYPES / FUNCTION / CATEGORY, same
argument meaning, IMAGE in ([B,H,W,3] float32) -> IMAGE out ([B,4H,4W,3] float32).

What changes is where the work happens: the ESRGAN / RRDBNet model, the reversal hook, the tiled feather blend and
the YCbCr recombination all run in libhdrvae.so on the GPU (vae_decode_hdr_b200/csrc/upscaler.cu).  There is no CPU
path: without a CUDA device or without the library the node raises.

Model loading stays what it is in ComfyUI (folder_paths + spandrel, :73-77); the node only reads
`descriptor.model.state_dict()`, `descriptor.scale` and `descriptor.architecture.name`.  Architectures other than
the 4x RRDBNet (nf 64, gc 32) are rejected by the library with an error naming the tensor that does not fit.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, Tuple

import torch

from .upscaler import UPSCALE_METHODS, HdrUpscalerEngine, reversal_for_architecture


def _model_filenames():
    try:
        import folder_paths  # provided by ComfyUI
        return folder_paths.get_filename_list("upscale_models")
    except Exception:
        return []


class HDRUpscaleWithModel:
    _engines: Dict[Any, HdrUpscalerEngine] = {}
    logger = logging.getLogger("HDRUpscaleWithModel")

    @classmethod
    def INPUT_TYPES(s):
        # hdr_upscale_with_model.py:59-66
        return {"required": {
            "image": ("IMAGE",),
            "model_name": (_model_filenames(),),
            "small_blur": ("BOOLEAN", {"default": False, "tooltip": "Apply small blur to avoid hot-pixels."}),
            "local_fix": ("BOOLEAN", {"default": False, "tooltip": "Apply local masking to suppress extreme hotspots in dark areas."}),
            "upscale_method": (list(UPSCALE_METHODS), {"default": "bislerp", "tooltip": "method used by the local_fix"}),
        }}

    RETURN_TYPES = ("IMAGE",)
    FUNCTION = "upscale"
    CATEGORY = "HDR/Upscale"

    # hdr_upscale_with_model.py:73-77 — unchanged: ComfyUI's model folder + spandrel's loader
    def _load_model_internal(self, model_name):
        import folder_paths
        from spandrel import ModelLoader
        return ModelLoader().load_from_file(folder_paths.get_full_path("upscale_models", model_name))

    @staticmethod
    def _compute_device(image: torch.Tensor) -> torch.device:
        if image.device.type == "cuda":
            return image.device
        if not torch.cuda.is_available():
            raise RuntimeError("HDRUpscaleWithModel (B200) needs a CUDA device; there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    @classmethod
    def _engine_for(cls, descriptor: Any, device: torch.device) -> HdrUpscalerEngine:
        model = descriptor.model
        params = list(model.parameters())
        key = (id(model), str(device), tuple((p.data_ptr(), p._version) for p in params[:4]))
        eng = cls._engines.get(key)
        if eng is None:
            for k in [k for k in cls._engines if k[0] == id(model) and k[1] == str(device)]:
                cls._engines.pop(k).close()
            eng = HdrUpscalerEngine(model.state_dict(), device)
            cls._engines[key] = eng
        return eng

    def upscale(self, image, model_name, small_blur=False, local_fix=False, upscale_method="bislerp") -> Tuple[torch.Tensor]:
        descriptor = self._load_model_internal(model_name)
        scale = getattr(descriptor, "scale", 4)
        if scale != 4:
            raise RuntimeError(f"HDRUpscaleWithModel (B200) implements 4x RRDBNet models; {model_name!r} has scale {scale}")
        device = self._compute_device(image)
        engine = self._engine_for(descriptor, device)
        reversal = reversal_for_architecture(descriptor.architecture.name)     # :266-279
        out = engine.upscale(image, reversal, bool(small_blur), bool(local_fix), upscale_method)
        if image.device.type != "cuda":
            out = out.to(image.device)
        return (out,)
