"""Import the UNMODIFIED reference (``/root/reference``) in the build container.

TEST INFRASTRUCTURE ONLY, and container-only: ``/root/reference`` does not
exist on the GPU box, so nothing in ``tests/ -m gpu``, ``smoke()`` or
``bench.py`` imports this module.  ``oracle/make_golden.py`` uses it to freeze
outputs of the reference's own functions into ``tests/golden/``.

What has to be shimmed:
* packages that are not installed (SimpleITK, timm, accelerate, ...) are
  replaced by ``MagicMock`` modules -- none of their arithmetic is reached by
  the functions we call, except timm, whose ``create_model`` is pointed at the
  timm-key-compatible restatement in ``oracle/convnext.py``;
* the reference snapshot has a circular import through
  ``spine_vision.training`` (``training/__init__.py:48`` ->
  ``models/generic.py:45`` -> ``registry.py:26`` -> ``trainers/__init__.py:11``
  -> ``trainers/classification.py:31``); it is bypassed by pre-seeding
  ``spine_vision.training`` with an empty package and
  ``spine_vision.training.registry`` with a stub whose ``register_model`` is
  the identity decorator.
"""

from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path
from unittest.mock import MagicMock

REFERENCE_ROOT = Path("/root/reference")

_ABSENT = [
    "SimpleITK", "accelerate", "torchmetrics", "iterstrat", "fitz", "openpyxl", "paddleocr",
    "vietocr", "rapidfuzz", "unidecode", "plotly", "trackio", "matplotlib", "seaborn", "pydicom",
]


class _MockFinder:
    """Meta-path finder: any (sub)module of an absent top-level package imports
    as a ``MagicMock`` package, so ``from openpyxl.reader.excel import X`` works."""

    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            from importlib.machinery import ModuleSpec

            return ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__path__ = []
        m.__name__ = spec.name
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        return None


def available() -> bool:
    return (REFERENCE_ROOT / "spine_vision").is_dir()


def install():
    """Make ``import spine_vision...`` work for the hot-path modules."""
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?) -- use tests/golden instead")
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    missing = []
    for name in _ABSENT:
        try:
            importlib.import_module(name)
        except Exception:
            missing.append(name)
    if missing and not any(isinstance(f, _MockFinder) for f in sys.meta_path):
        sys.meta_path.append(_MockFinder(missing))

    # timm shim -> oracle ConvNeXt with timm's key names
    from oracle.convnext import ConvNeXt

    timm = types.ModuleType("timm")

    def create_model(name: str, pretrained: bool = False, num_classes: int = 0, **kw):
        assert num_classes == 0 and not pretrained, "shim supports the hot path's call only"
        variant = name.split(".")[0].replace("convnext_", "")
        return ConvNeXt(variant)

    timm.create_model = create_model  # type: ignore[attr-defined]
    sys.modules["timm"] = timm

    # circular-import bypass
    pkg = types.ModuleType("spine_vision.training")
    pkg.__path__ = [str(REFERENCE_ROOT / "spine_vision" / "training")]  # type: ignore[attr-defined]
    sys.modules["spine_vision.training"] = pkg
    reg = types.ModuleType("spine_vision.training.registry")

    def register_model(*_a, **_k):
        return lambda cls: cls

    reg.register_model = register_model  # type: ignore[attr-defined]
    reg.register_trainer = register_model  # type: ignore[attr-defined]
    reg.register_metrics = register_model  # type: ignore[attr-defined]
    reg.ModelRegistry = MagicMock()  # type: ignore[attr-defined]
    sys.modules["spine_vision.training.registry"] = reg


def load():
    """Return a namespace with the reference's own hot-path callables."""
    install()
    from spine_vision.io import normalize_to_uint8  # io/__init__.py:15
    from spine_vision.datasets.classification import cropping  # cropping.py

    ns = types.SimpleNamespace(
        normalize_to_uint8=normalize_to_uint8,
        resize_with_padding=cropping.resize_with_padding,
        mm_to_pixels=cropping.mm_to_pixels,
        crop_region_horizontal=cropping.crop_region_horizontal,
        crop_region_rotated=cropping.crop_region_rotated,
        get_rotation_angles=cropping.get_rotation_angles,
        CropContext=cropping.CropContext,
        predict_ivd_locations=cropping.predict_ivd_locations,
        load_localization_model=cropping.load_localization_model,
        cropping=cropping,
    )
    try:
        from spine_vision.training.models.generic import CoordinateRegressor

        ns.CoordinateRegressor = CoordinateRegressor
    except Exception as e:  # pragma: no cover - diagnostic only
        ns.CoordinateRegressor = None
        ns.coordinate_regressor_error = repr(e)
    try:
        from spine_vision.training.datasets.classification import construct_3channel

        ns.construct_3channel = construct_3channel
    except Exception as e:  # pragma: no cover
        ns.construct_3channel = None
        ns.construct_3channel_error = repr(e)
    return ns
