"""Implementation switch and registry hook (SURVEY.md section 5 "Config / flags", Appendix B).

The reference builds its modules BY NAME: `encoder_factory("croco", **kw)` / `feature_returner_encoder_factory` look a class up
in `ENCODER_CONFIGS` (uniception/models/encoders/__init__.py:37-140) and downstream models index
`INFO_SHARING_CLASSES["cross_attention"]` (uniception/models/info_sharing/__init__.py:23-37).  Two ways to get the B200 path
behind those names:

  * explicit, per module:   registry.build_encoder("croco", implementation="b200", **kw)
                            registry.info_sharing_class("cross_attention", ifr=False, implementation="reference")
  * global hook:            registry.install("b200")   # patches the reference's own registries in place, so code that calls
                            ...                        # uniception.models.encoders.encoder_factory("croco", ...) unchanged
                            registry.install("reference")   # restores the reference's classes (== uninstall())
    or `with registry.implementation("b200"): ...`

Only the names this package implements are swapped (croco; cross_attention / alternating_attention / global_attention /
diff_cross_attention); every other encoder keeps the reference's class.  Nothing here imports `uniception` unless a
"reference" implementation or the hook is asked for -- importing uniception_b200 never pulls the reference in.
"""
from __future__ import annotations

import contextlib
import importlib
from typing import Dict, Tuple

from . import encoders as _enc
from . import info_sharing as _info

IMPLEMENTATIONS = ("b200", "reference")
_ENCODER_NAMES = ("croco",)
_INFO_NAMES = tuple(_info.INFO_SHARING_CLASSES.keys())
_saved: Dict[Tuple[str, str], object] = {}  # ("enc" | "info", registry name) -> the reference's own entry
_current = "reference"  # what the REFERENCE's registries currently resolve to


def _reference_modules():
    try:
        enc = importlib.import_module("uniception.models.encoders")
        info = importlib.import_module("uniception.models.info_sharing")
    except ImportError as exc:  # the reference is optional: the drop-in classes work without it
        raise ImportError("the reference package `uniception` is not importable; only implementation='b200' is available") from exc
    return enc, info


def _check(impl: str) -> None:
    if impl not in IMPLEMENTATIONS:
        raise ValueError(f"Unknown implementation: {impl}. Valid options: {IMPLEMENTATIONS}")


def encoder_class(encoder_str: str, *, feature_returner: bool = False, implementation: str = "b200"):
    """The class registered under `encoder_str` in the chosen implementation (ValueError for unknown names, like
    encoders/__init__.py:109-112)."""
    _check(implementation)
    key = "intermediate_feature_returner_class" if feature_returner else "class"
    if implementation == "b200":
        if encoder_str not in _enc.ENCODER_CONFIGS:
            raise ValueError(f"Unknown encoder: {encoder_str}. The B200 path implements {list(_enc.ENCODER_CONFIGS)}")
        return _enc.ENCODER_CONFIGS[encoder_str][key]
    enc, _ = _reference_modules()
    if encoder_str not in enc.ENCODER_CONFIGS:
        raise ValueError(f"Unknown encoder: {encoder_str}.")
    cfg = _saved.get(("enc", encoder_str), enc.ENCODER_CONFIGS[encoder_str])
    return cfg[key]


def build_encoder(encoder_str: str, *, implementation: str = "b200", feature_returner: bool = False, **kwargs):
    """`encoder_factory` / `feature_returner_encoder_factory` with an explicit implementation."""
    return encoder_class(encoder_str, feature_returner=feature_returner, implementation=implementation)(**kwargs)


def info_sharing_class(name: str, *, ifr: bool = False, implementation: str = "b200"):
    """`INFO_SHARING_CLASSES[name][int(ifr)]` of the chosen implementation."""
    _check(implementation)
    if implementation == "b200":
        if name not in _info.INFO_SHARING_CLASSES:
            raise ValueError(f"Unknown info-sharing type: {name}. The B200 path implements {list(_info.INFO_SHARING_CLASSES)}")
        return _info.INFO_SHARING_CLASSES[name][int(ifr)]
    _, info = _reference_modules()
    if name not in info.INFO_SHARING_CLASSES:
        raise ValueError(f"Unknown info-sharing type: {name}.")
    return _saved.get(("info", name), info.INFO_SHARING_CLASSES[name])[int(ifr)]


def install(implementation: str = "b200") -> str:
    """Point the REFERENCE's registries at the chosen implementation (in place).  Returns the previous setting."""
    global _current
    _check(implementation)
    enc, info = _reference_modules()
    prev = _current
    if implementation == "b200":
        for n in _ENCODER_NAMES:
            if n in enc.ENCODER_CONFIGS and ("enc", n) not in _saved:
                _saved[("enc", n)] = dict(enc.ENCODER_CONFIGS[n])
                enc.ENCODER_CONFIGS[n] = {**enc.ENCODER_CONFIGS[n], **_enc.ENCODER_CONFIGS[n]}
        for n in _INFO_NAMES:
            if n in info.INFO_SHARING_CLASSES and ("info", n) not in _saved:
                _saved[("info", n)] = info.INFO_SHARING_CLASSES[n]
                info.INFO_SHARING_CLASSES[n] = _info.INFO_SHARING_CLASSES[n]
    else:
        for (kind, n), old in list(_saved.items()):
            if kind == "enc":
                enc.ENCODER_CONFIGS[n] = old
            else:
                info.INFO_SHARING_CLASSES[n] = old
            del _saved[(kind, n)]
    _current = implementation
    return prev


def uninstall() -> None:
    """Restore the reference's own classes in its registries."""
    if _saved:
        install("reference")


def current() -> str:
    return _current


@contextlib.contextmanager
def implementation(impl: str):
    """`with registry.implementation("b200"): model = SomeReferenceFactory(...)`"""
    prev = install(impl)
    try:
        yield
    finally:
        install(prev)
