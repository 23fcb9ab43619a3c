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
