"""Import the reference's own ``ObjectDetection`` head, unmodified.

TEST INFRASTRUCTURE ONLY.  Works only where the reference source tree exists
(``/root/reference`` in the authoring container; it does not exist on the GPU
box, so nothing marked ``gpu`` and neither ``smoke()`` nor ``bench.py`` may call
this).  Used by ``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and
by the ``not gpu`` tests that validate ``oracle/torch_restatement.py`` against
the real thing.

``import sihl`` cannot work here: ``torchmetrics`` / ``lightning`` are not
installed and ``sihl/__init__.py`` asks for package metadata.  The recipe
(SURVEY.md §8c): stub ``torchmetrics`` in ``sys.modules`` and load
``src/sihl/heads/object_detection.py`` straight from its file.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SIHL_REFERENCE_ROOT", "/root/reference")
_OD_FILE = os.path.join(REFERENCE_ROOT, "src", "sihl", "heads", "object_detection.py")


def available() -> bool:
    return os.path.isfile(_OD_FILE)


def _stub_torchmetrics() -> None:
    if "torchmetrics" in sys.modules:
        return

    class _Unavailable:
        def __init__(self, *a, **k):
            raise RuntimeError("torchmetrics is stubbed in this container")

    tm = types.ModuleType("torchmetrics")
    tm.MeanMetric = _Unavailable
    tm.Metric = _Unavailable
    det = types.ModuleType("torchmetrics.detection")
    mean_ap = types.ModuleType("torchmetrics.detection.mean_ap")
    mean_ap.MeanAveragePrecision = _Unavailable
    tm.detection, det.mean_ap = det, mean_ap
    sys.modules.update({"torchmetrics": tm, "torchmetrics.detection": det,
                        "torchmetrics.detection.mean_ap": mean_ap})


_module = None


def load_module():
    """The reference module object for ``sihl/heads/object_detection.py``."""
    global _module
    if _module is None:
        if not available():
            raise FileNotFoundError(f"reference source not found at {_OD_FILE}")
        _stub_torchmetrics()
        spec = importlib.util.spec_from_file_location("_sihl_reference_object_detection", _OD_FILE)
        _module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_module)
    return _module


def ObjectDetection():
    """The reference class itself (call it to construct a head)."""
    return load_module().ObjectDetection


_quad_module = None


def QuadrilateralDetection():
    """The reference ``QuadrilateralDetection`` class, loaded straight from its file."""
    global _quad_module
    if _quad_module is None:
        path = os.path.join(REFERENCE_ROOT, "src", "sihl", "heads", "quadrilateral_detection.py")
        if not os.path.isfile(path):
            raise FileNotFoundError(f"reference source not found at {path}")
        _stub_torchmetrics()
        spec = importlib.util.spec_from_file_location("_sihl_reference_quadrilateral_detection", path)
        _quad_module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_quad_module)
    return _quad_module.QuadrilateralDetection
