"""Import the UNMODIFIED reference node from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may call :func:`load_reference_node`; it is used by
``make_golden.py`` (to produce tests/golden/*.npz) and by the not-gpu tests that
re-validate the restated oracle when the reference tree is present.

The reference file's top-level ``from kornia.core import ImageModule, Tensor``
(hdr_vae_decode.py:15-16) is unused but makes the import fail when kornia is
absent; a 4-line ``sys.modules`` stub makes the file
importable without touching it.
"""
from __future__ import annotations

import importlib.util
import logging
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("HDRVAE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hdr_vae_decode.py"))


def _stub_kornia():
    if "kornia" in sys.modules:
        return
    try:
        import kornia  # noqa: F401
        return
    except Exception:
        pass
    k = types.ModuleType("kornia")
    kc = types.ModuleType("kornia.core")
    kc.ImageModule = torch.nn.Module
    kc.Tensor = torch.Tensor
    k.core = kc
    sys.modules["kornia"] = k
    sys.modules["kornia.core"] = kc


def load_reference_module(name: str = "hdr_vae_decode"):
    if not reference_available():
        raise FileNotFoundError(f"reference tree not present at {REFERENCE_ROOT}")
    _stub_kornia()
    modname = f"_hdrvae_reference_{name}"
    if modname in sys.modules:
        return sys.modules[modname]
    root_level = logging.getLogger().level
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    # the reference calls logging.basicConfig(level=INFO) at import (hdr_vae_decode.py:19);
    # keep the test output readable.
    logging.getLogger().setLevel(root_level if root_level else logging.WARNING)
    logging.getLogger(modname).setLevel(logging.WARNING)
    return mod


def load_reference_node():
    """Fresh ``HDRVAEDecode`` instance of the unmodified reference."""
    mod = load_reference_module("hdr_vae_decode")
    node = mod.HDRVAEDecode()
    node.logger.setLevel(logging.WARNING)
    return node


def _stub_upscaler_deps(descriptor):
    """Stand-ins for the packages hdr_upscale_with_model.py imports that are not installed here (folder_paths,
    comfy, spandrel, kornia.color / kornia.filters): the restated algorithms of oracle.upscaler_oracle, wired so that
    the UNMODIFIED reference node runs on CPU.  `descriptor` is what `ModelLoader().load_from_file` returns."""
    from oracle import upscaler_oracle as uo
    _stub_kornia()
    k = sys.modules["kornia"]
    if not hasattr(k, "color") or not hasattr(getattr(k, "color"), "rgb_to_ycbcr"):
        kcol, kfil = types.ModuleType("kornia.color"), types.ModuleType("kornia.filters")
        kcol.rgb_to_ycbcr = uo.rgb_to_ycbcr
        kfil.median_blur = uo.median_blur
        k.color, k.filters = kcol, kfil
        sys.modules["kornia.color"], sys.modules["kornia.filters"] = kcol, kfil
    fp = types.ModuleType("folder_paths")
    fp.get_filename_list = lambda kind: ["fake_esrgan.pth"]
    fp.get_full_path = lambda kind, name: name
    sys.modules["folder_paths"] = fp
    comfy = types.ModuleType("comfy")
    mm, cu = types.ModuleType("comfy.model_management"), types.ModuleType("comfy.utils")
    mm.get_torch_device = lambda: torch.device("cpu")
    mm.module_size = lambda m: 0
    mm.free_memory = lambda *a, **k: None
    mm.OOM_EXCEPTION = torch.OutOfMemoryError if hasattr(torch, "OutOfMemoryError") else MemoryError

    class _PBar:
        def __init__(self, total):
            self.total = total

        def update(self, n):
            pass
    cu.ProgressBar = _PBar
    cu.tiled_scale = uo.tiled_scale
    cu.get_tiled_scale_steps = uo.get_tiled_scale_steps
    cu.common_upscale = uo.common_upscale
    comfy.model_management, comfy.utils = mm, cu
    sys.modules["comfy"], sys.modules["comfy.model_management"], sys.modules["comfy.utils"] = comfy, mm, cu
    sp = types.ModuleType("spandrel")

    class ModelLoader:
        def load_from_file(self, path):
            return sp._descriptor
    sp.ModelLoader = ModelLoader
    sp.ImageModelDescriptor = type(descriptor)
    sp._descriptor = descriptor
    sys.modules["spandrel"] = sp


def load_reference_upscaler(descriptor):
    """Fresh `HDRUpscaleWithModel` of the unmodified reference whose model loader returns `descriptor`."""
    _stub_upscaler_deps(descriptor)
    sys.modules.pop("_hdrvae_reference_hdr_upscale_with_model", None)
    mod = load_reference_module("hdr_upscale_with_model")
    sys.modules["spandrel"]._descriptor = descriptor
    return mod.HDRUpscaleWithModel()
