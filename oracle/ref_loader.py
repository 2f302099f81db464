"""Import the reference's own hot-path files BY PATH (authoring container only).

``/root/reference`` does not exist on the GPU box, so nothing in the ``-m gpu``
tests, ``smoke()`` or ``bench.py`` may call this.  It is used by
``oracle/gen_golden.py`` (to write ``tests/golden/*.npz``) and by the CPU tests
that validate the restatement against the live reference when it is present.

The reference cannot be imported as a package here (``core/models/__init__.py``
pulls in mmcv; ``core/utils/utility.py:15`` imports matplotlib), so the three
files on the path are loaded one by one with stub modules for what is absent.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200SEG_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "core/utils/utility.py"))


def _load(name: str, relpath: str):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.cm = types.SimpleNamespace(Reds=None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


_cache = {}


def load_reference():
    """Returns a namespace with the reference's own classes/functions on the path."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _stub_matplotlib()
    classifier = _load("_ref_aspp_classifier", "core/models/classifiers/aspp/classifier.py")
    discriminator = _load("_ref_discriminator", "core/models/discriminator.py")
    utility = _load("_ref_utility", "core/utils/utility.py")
    ns = types.SimpleNamespace(
        ASPP_Classifier_V2=classifier.ASPP_Classifier_V2,
        PixelDiscriminator=discriminator.PixelDiscriminator,
        soft_label_cross_entropy=utility.soft_label_cross_entropy,
        inference=utility.inference,
        multi_scale_inference=utility.multi_scale_inference,
        get_color_palette=utility.get_color_palette,
        adjust_learning_rate=_load("_ref_adapt_lr", "core/utils/adapt_lr.py").adjust_learning_rate,
        intersectionAndUnion=utility.intersectionAndUnion,
        confusion_matrix=utility.confusion_matrix,
        AverageMeter=utility.AverageMeter,
    )
    _cache["ns"] = ns
    return ns
