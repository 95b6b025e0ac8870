"""Import the reference's own code, unmodified.

TEST INFRASTRUCTURE ONLY (``tests/``, ``oracle/make_golden.py`` and the reference legs of ``bench.py``; nothing under
``sihl_b200/`` imports this).

Where the reference package is looked for, in order:
  1. ``$SIHL_REFERENCE_ROOT/src/sihl`` (default ``/root/reference/src/sihl``) — the authoring container;
  2. ``oracle/_ref/sihl_src/sihl`` — the byte-for-byte staged copy made by ``oracle/stage_reference.py``
     (git-ignored, travels to the GPU box with the ``gpurun`` snapshot): this is what makes the *real* reference the
     same-device oracle on the B200 box;
  3. ``baseline/_ref/sihl`` — a ``pip install --target`` of the reference, if a driver put one there.

``import sihl`` itself cannot work: ``sihl/__init__.py`` asks for package metadata, ``torchmetrics`` / ``lightning`` are
not installed and ``sihl/heads/__init__.py`` imports all 14 heads.  The recipe: stub ``torchmetrics`` in
``sys.modules`` (a working ``MeanMetric``, a recording ``MeanAveragePrecision``), register ``sihl`` and ``sihl.heads`` as
bare namespace packages pointing at the reference directories (their ``__init__`` files are skipped, every other file
is executed as is), then import ``sihl.heads.object_detection`` & co. normally.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SIHL_REFERENCE_ROOT", "/root/reference")
_CANDIDATES = (
    ("reference", os.path.join(REFERENCE_ROOT, "src", "sihl")),
    ("oracle/_ref", os.path.join(HERE, "_ref", "sihl_src", "sihl")),
    ("baseline/_ref", os.path.join(os.path.dirname(HERE), "baseline", "_ref", "sihl")),
)


def package_dir():
    """``(kind, directory)`` of the first place the reference package is found, or ``(None, None)``."""
    for kind, path in _CANDIDATES:
        if os.path.isfile(os.path.join(path, "heads", "object_detection.py")):
            return kind, path
    return None, None


def available() -> bool:
    return package_dir()[1] is not None


def kind() -> str:
    """Where the reference code comes from: "reference" (/root/reference), "oracle/_ref" (staged copy) or
    "baseline/_ref"; ``bench.py`` reports it as ``cpu_baseline.kind``."""
    return package_dir()[0] or "unavailable"


# --------------------------------------------------------------------------- torchmetrics stand-ins
class _MeanMetric:
    """``torchmetrics.MeanMetric(nan_strategy="ignore")`` as far as the head uses it (ref :220,:239,:249)."""

    def __init__(self, nan_strategy="warn", **kwargs):
        self.nan_strategy, self.total, self.count = nan_strategy, 0.0, 0

    def to(self, device):
        return self

    def update(self, value):
        v = float(value.detach()) if hasattr(value, "detach") else float(value)
        if v == v:
            self.total, self.count = self.total + v, self.count + 1

    def compute(self):
        import torch
        return torch.tensor(self.total / self.count if self.count else float("nan"))

    def reset(self):
        self.total, self.count = 0.0, 0


class _RecordingMAP:
    """Stand-in for ``MeanAveragePrecision``: records what the head hands to ``update`` (ref :230-237) so a test can
    compare it with what the replacement hands over; ``compute`` returns the number of recorded images only —
    torchmetrics and its ``faster_coco_eval`` backend are not installed, the metric itself is out of reach here."""

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs, self.preds, self.targets = args, kwargs, [], []

    def to(self, device):
        return self

    def update(self, preds, targets):
        self.preds.extend(preds)
        self.targets.extend(targets)

    def compute(self):
        import torch
        return {"recorded_images": torch.tensor(len(self.preds))}


class _Metric:
    """Base class the reference's own metric helpers derive from (``utils/f1.py``, ``utils/pck.py``); never instantiated
    on this path."""

    def __init__(self, *a, **k):
        raise RuntimeError("torchmetrics is stubbed in this container")


def _stub_torchmetrics() -> None:
    if "torchmetrics" in sys.modules:
        return
    tm = types.ModuleType("torchmetrics")
    tm.MeanMetric = _MeanMetric
    tm.Metric = _Metric
    tm.__sihl_b200_stub__ = True
    det = types.ModuleType("torchmetrics.detection")
    mean_ap = types.ModuleType("torchmetrics.detection.mean_ap")
    mean_ap.MeanAveragePrecision = _RecordingMAP
    tm.detection, det.mean_ap = det, mean_ap
    sys.modules.update({"torchmetrics": tm, "torchmetrics.detection": det,
                        "torchmetrics.detection.mean_ap": mean_ap})


# --------------------------------------------------------------------------- the package
_registered = False


def _register_namespace() -> str:
    """Make ``sihl.*`` importable from the reference directory without running ``sihl/__init__.py`` and
    ``sihl/heads/__init__.py`` (metadata lookup / 14 heads + lightning).  Other sub-packages (``sihl.layers``,
    ``sihl.utils``) import normally, with their own ``__init__``."""
    global _registered
    _, path = package_dir()
    if path is None:
        raise FileNotFoundError("reference source not found: looked in " + ", ".join(p for _, p in _CANDIDATES) +
                                " (run `python -m oracle.stage_reference` in the authoring container)")
    if not _registered:
        _stub_torchmetrics()
        for name, sub in (("sihl", path), ("sihl.heads", os.path.join(path, "heads"))):
            if name in sys.modules and not getattr(sys.modules[name], "__sihl_b200_namespace__", False):
                raise RuntimeError(f"a real `{name}` package is already imported; the loader would shadow it")
            mod = types.ModuleType(name)
            mod.__path__ = [sub]
            mod.__sihl_b200_namespace__ = True
            sys.modules[name] = mod
        sys.modules["sihl"].heads = sys.modules["sihl.heads"]
        _registered = True
    return path


def load(module: str):
    """Import ``sihl.<module>`` from the reference tree, e.g. ``load("heads.object_detection")``."""
    _register_namespace()
    return importlib.import_module("sihl." + module)


def load_module():
    """The reference module object for ``sihl/heads/object_detection.py``."""
    return load("heads.object_detection")


def ObjectDetection():
    """The reference class itself (call it to construct a head)."""
    return load_module().ObjectDetection


def QuadrilateralDetection():
    """The reference ``QuadrilateralDetection`` class."""
    return load("heads.quadrilateral_detection").QuadrilateralDetection


def SihlModel():
    """ref: src/sihl/sihl_model.py:6-25 — the backbone -> neck -> heads container the Lightning module drives."""
    return load("sihl_model").SihlModel


def TorchvisionBackbone():
    """ref: src/sihl/torchvision_backbone.py:101 (config[0]: resnet18, no pretrained weights)."""
    return load("torchvision_backbone").TorchvisionBackbone


def FPN():
    """ref: src/sihl/layers/fpn.py:8."""
    return load("layers.fpn").FPN
