"""HDRVAEDecode — the ComfyUI node, same surface as the reference (hdr_vae_decode.py:23-60),
computing on libhdrvae.so (sm_100a) instead of PyTorch eager.

Kept verbatim from the reference: INPUT_TYPES (required samples/vae, optional hdr_mode enum in
the code's order with default "mathematical_recovery", conservative_ev_multiplier FLOAT 0.1-10),
RETURN_TYPES ("IMAGE",), RETURN_NAMES ("image",), FUNCTION "simple_hdr_decode", CATEGORY "latent",
and the float32 [B,H,W,3] contiguous output (:195,:354).

Differences, all documented in DESIGN.md:
  * one decoder pass instead of two + a third conv_out;
  * README-era mode names "moderate"/"aggressive" are accepted as aliases;
  * NORMALIZATION_FUNCTION is recomputed per call (the reference's sticky state, SURVEY.md §0.9, is a bug);
  * the non-deterministic bypass fallback is replaced by a deterministic rule: the
    intelligent result is always returned and ``last_stats['accepted']`` tells whether the reference
    would have gone to its bypass.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, Optional, Tuple

import torch

from .engine import DEFAULT_MODE, HDR_MODES, HdrVaeEngine, quantiles

logger = logging.getLogger(__name__)


def _decoder_of(vae: Any):
    try:
        return vae.first_stage_model.decoder          # hdr_vae_decode.py:842
    except AttributeError as e:
        raise RuntimeError("HDRVAEDecode: vae.first_stage_model.decoder not found (not a Flux/SD-style VAE)") from e


def _weights_key(decoder) -> Tuple:
    ps = list(decoder.parameters())
    return (id(decoder), len(ps), tuple((p.data_ptr(), p._version) for p in ps[:4] + ps[-4:]))


class HDRVAEDecode:
    """HDR VAE Decode (B200-native).  Drop-in for the reference node class of the same name."""

    _engines: Dict[Tuple, HdrVaeEngine] = {}

    def __init__(self):
        self.logger = logger
        self.NORMALIZATION_FUNCTION = str()
        self.last_stats: Optional[Dict] = None
        self.profile_quantiles = False      # True: last_stats["out_quantiles"] = {q: value} of the output image

    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {
                "samples": ("LATENT",),
                "vae": ("VAE",),
            },
            "optional": {
                "hdr_mode": (list(HDR_MODES),
                             {"default": DEFAULT_MODE,
                              "tooltip": "conservative: Gentle conservative_ev_multiplier expansion, safest for general use \n "
                                         "exposure: Natural exposure-based HDR for compositing workflows \n "
                                         "mathematical_recovery: Full mathematical recovery, maximum range"}),
                "conservative_ev_multiplier": ("FLOAT", {"default": 1.0, "min": 0.1, "max": 10.0, "step": 0.1,
                                                         "tooltip": "Expansion multiplier for the conservative mode."}),
            }
        }

    RETURN_TYPES = ("IMAGE",)
    RETURN_NAMES = ("image",)
    FUNCTION = "simple_hdr_decode"
    CATEGORY = "latent"

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _compute_device(vae: Any, latent: torch.Tensor) -> torch.device:
        for cand in (getattr(vae, "device", None), latent.device):
            if cand is not None and torch.device(cand).type == "cuda":
                return torch.device(cand)
        if not torch.cuda.is_available():
            raise RuntimeError("HDRVAEDecode (B200) needs a CUDA device; there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    @classmethod
    def _engine_for(cls, vae: Any, device: torch.device) -> HdrVaeEngine:
        decoder = _decoder_of(vae)
        key = (_weights_key(decoder), str(device))
        eng = cls._engines.get(key)
        if eng is None:
            for k in [k for k in cls._engines if k[0][0] == id(decoder) and k[1] == str(device)]:
                cls._engines.pop(k).close()           # weights changed in place: repack
            eng = HdrVaeEngine(decoder.state_dict(), device)
            cls._engines[key] = eng
        return eng

    @classmethod
    def adopt_engine(cls, vae: Any, device, engine: HdrVaeEngine) -> None:
        """Register an already-built engine for this vae's decoder weights (avoids a second repack)."""
        cls._engines[(_weights_key(_decoder_of(vae)), str(torch.device(device)))] = engine

    def simple_hdr_decode(
        self,
        samples: Dict[str, torch.Tensor],
        vae: Any,
        hdr_mode: str = DEFAULT_MODE,
        conservative_ev_multiplier: float = 1.0,
    ) -> Tuple[torch.Tensor]:
        latent = samples["samples"]                    # hdr_vae_decode.py:78
        device = self._compute_device(vae, latent)
        engine = self._engine_for(vae, device)
        image, stats = engine.decode(latent, hdr_mode, conservative_ev_multiplier, want_stats=True)
        if self.profile_quantiles:
            # statistical profiling beyond the reference's min/max/mean: exact quantiles by GPU radix select
            qs = (0.01, 0.5, 0.99, 0.999)
            stats["out_quantiles"] = dict(zip(qs, quantiles(image, qs)))
        self.last_stats = stats
        self.NORMALIZATION_FUNCTION = {0: "", 1: "SIGMOID", 2: "TANH"}[stats["norm_function"]]
        if not stats["accepted"]:
            self.logger.warning("HDRVAEDecode: no value above 1.0 in the intelligent result; the reference would "
                                "fall back to its non-deterministic bypass decode here (returning the deterministic "
                                "intelligent result)")
        self.logger.debug("OUTPUT: range=[%.3f, %.3f], HDR pixels: %d, Negative pixels: %d", stats["out_min"],
                          stats["out_max"], stats["hdr_pixels"], stats["negative_pixels"])
        out_dev = getattr(vae, "output_device", None)
        if out_dev is None:
            out_dev = latent.device
        out_dev = torch.device(out_dev)
        if out_dev != image.device:
            if out_dev.type == "cpu":
                # pinned staging (torch's caching host allocator recycles it): D2H at link speed
                host = torch.empty(image.shape, dtype=image.dtype, pin_memory=True)
                host.copy_(image, non_blocking=True)
                torch.cuda.current_stream(image.device).synchronize()
                image = host
            else:
                image = image.to(out_dev)
        return (image,)
