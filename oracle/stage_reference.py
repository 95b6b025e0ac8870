"""Stage the reference's Python package next to the oracle so that it can travel to the GPU box.

TEST INFRASTRUCTURE ONLY.  The reference (jonregef/sihl) is pure Python: there is nothing to compile, but
``/root/reference`` does not exist on the GPU box, so the reference cannot be the same-device oracle there unless
its source files ride along.  This recipe copies the ``*.py`` files of ``/root/reference/src/sihl`` — unmodified,
byte for byte — into ``oracle/_ref/sihl_src/sihl/``.  ``oracle/_ref/`` is git-ignored (nothing of the reference ever
enters the history or the product package) but it is part of the ``gpurun`` snapshot, exactly like a compiled
``oracle/_ref/*.so`` would be.  ``oracle/ref_loader.py`` searches the staged copy after ``/root/reference``.

    python -m oracle.stage_reference        # also run by __graft_entry__.build() when /root/reference exists

The staged tree is only ever *imported by the checker*: ``tests/`` (GPU parity of the drop-in head under the
reference's own ``SihlModel`` / ``TorchvisionBackbone`` / ``FPN`` caller) and ``bench.py``'s reference legs
(``--impl reference``, ``cpu_baseline``, ``gpu_eager_reference``).  Nothing under ``sihl_b200/`` reads it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE_ROOT = os.environ.get("SIHL_REFERENCE_ROOT", "/root/reference")
STAGE_DIR = os.path.join(HERE, "_ref", "sihl_src")
MANIFEST = os.path.join(STAGE_DIR, "MANIFEST.json")


def source_available() -> bool:
    return os.path.isfile(os.path.join(SOURCE_ROOT, "src", "sihl", "heads", "object_detection.py"))


def staged_available() -> bool:
    return os.path.isfile(os.path.join(STAGE_DIR, "sihl", "heads", "object_detection.py"))


def stage(verbose: bool = False) -> str:
    """Copy the reference package (``*.py`` only) into ``oracle/_ref/sihl_src/sihl``; returns the staged root.
    Idempotent: files whose content is unchanged are left alone; a manifest records the sha256 of every file."""
    if not source_available():
        raise FileNotFoundError(f"reference source tree not found under {SOURCE_ROOT}")
    src_pkg = os.path.join(SOURCE_ROOT, "src", "sihl")
    dst_pkg = os.path.join(STAGE_DIR, "sihl")
    manifest = {}
    for dirpath, _, files in os.walk(src_pkg):
        rel_dir = os.path.relpath(dirpath, src_pkg)
        for name in sorted(files):
            if not name.endswith(".py"):
                continue
            src = os.path.join(dirpath, name)
            dst = os.path.normpath(os.path.join(dst_pkg, rel_dir, name))
            with open(src, "rb") as fh:
                data = fh.read()
            manifest[os.path.normpath(os.path.join(rel_dir, name))] = hashlib.sha256(data).hexdigest()
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            try:
                with open(dst, "rb") as fh:
                    if fh.read() == data:
                        continue
            except OSError:
                pass
            shutil.copyfile(src, dst)
            if verbose:
                print("staged", os.path.relpath(dst, HERE))
    version = "unknown"
    try:
        with open(os.path.join(SOURCE_ROOT, "pyproject.toml")) as fh:
            for line in fh:
                if line.strip().startswith("version"):
                    version = line.split("=", 1)[1].strip().strip('"')
                    break
    except OSError:
        pass
    with open(MANIFEST, "w") as fh:
        json.dump({"source": src_pkg, "version": version, "files": manifest}, fh, indent=1, sort_keys=True)
    return STAGE_DIR


if __name__ == "__main__":
    print(stage(verbose="-v" in sys.argv))
