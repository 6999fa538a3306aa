"""TEST INFRASTRUCTURE ONLY -- makes the real reference importable in the build container.

`/root/reference` (castacks/UniCeption @ 802ebc17) is pure PyTorch but its
`uniception/models/encoders/__init__.py:29` eagerly imports the Perception-Encoder
wrapper, which needs `timm.layers.DropPath` (libs/perception_encoder/vision_encoder/pe.py:14).
`timm` is not installed and there is no network, so a stub module is registered first.
This file is only used by `oracle/make_golden.py` (run once, here) and by CPU tests that
are skipped when the reference tree is absent (it does not exist on the GPU box).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("UC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "uniception"))


def import_reference():
    """Return the imported `uniception` package of the reference (or raise ImportError)."""
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    if "timm" not in sys.modules:
        import torch.nn as nn

        class DropPath(nn.Identity):  # drop_path=0 everywhere on the DUSt3R path
            def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
                super().__init__()

        timm = types.ModuleType("timm")
        layers = types.ModuleType("timm.layers")
        layers.DropPath = DropPath
        timm.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.layers"] = layers
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import uniception  # noqa: F401

    return uniception
