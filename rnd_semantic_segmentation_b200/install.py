"""Patch the B200 hot path into an importable copy of the reference tree.

    import rnd_semantic_segmentation_b200 as b200seg
    b200seg.install()          # before train_src.py / train_adv.py / test.py build their models

After this, ``core.models.build.build_classifier`` / ``build_adversarial_discriminator`` and
``core.utils.utility.{soft_label_cross_entropy,inference,intersectionAndUnionGPU,confusion_matrix,AverageMeter}``
resolve to this package's implementations; the scripts themselves stay unmodified.
"""
from __future__ import annotations

import importlib
import sys

from . import build as _build
from . import utility as _utility

_BUILD_NAMES = ("build_classifier", "build_adversarial_discriminator")
_UTIL_NAMES = ("soft_label_cross_entropy", "inference", "intersectionAndUnionGPU", "confusion_matrix", "AverageMeter")


def install(strict: bool = False):
    """Returns the list of 'module.attr' names that were patched."""
    patched = []
    targets = (("core.models.build", _build, _BUILD_NAMES), ("core.models", _build, _BUILD_NAMES),
               ("core.utils.utility", _utility, _UTIL_NAMES))
    for modname, src, names in targets:
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                if strict:
                    raise
                continue
        for n in names:
            setattr(mod, n, getattr(src, n))
            patched.append(f"{modname}.{n}")
    # modules that did `from core.utils.utility import X` / `from core.models import build_*` before install()
    for mod in list(sys.modules.values()):
        name = getattr(mod, "__name__", "")
        if not name.startswith("core.") or name in ("core.models.build", "core.models", "core.utils.utility"):
            continue
        for src, names in ((_build, _BUILD_NAMES), (_utility, _UTIL_NAMES)):
            for n in names:
                if n in getattr(mod, "__dict__", {}):
                    setattr(mod, n, getattr(src, n))
                    patched.append(f"{name}.{n}")
    return patched
