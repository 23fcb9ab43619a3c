"""vae_decode_hdr_b200 — B200-native (sm_100a) drop-in for the nodes of netocg/vae-decode-hdr.
Registered exactly like the reference package (__init__.py:43-53): the same three node keys and display names."""
from .hdr_upscale_with_model import HDRUpscaleWithModel
from .hdr_vae_decode import HDRVAEDecode
from .linear_exr_export import LinearEXRExport

NODE_CLASS_MAPPINGS = {
    "HDRVAEDecode": HDRVAEDecode,
    "LinearEXRExport": LinearEXRExport,
    "HDRUpscaleWithModel": HDRUpscaleWithModel,
}

NODE_DISPLAY_NAME_MAPPINGS = {
    "HDRVAEDecode": "HDR VAE Decode",
    "LinearEXRExport": "Linear EXR Export",
    "HDRUpscaleWithModel": "HDR Upscale with Model",
}

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS", "HDRVAEDecode", "LinearEXRExport", "HDRUpscaleWithModel"]
