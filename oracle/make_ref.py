#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference, made importable on the GPU box -- TEST INFRASTRUCTURE ONLY.

The reference (mnhrk15/integrated_path_planning) is pure Python; `/root/reference` exists in the build
container only.  This recipe copies its package tree `src/` and the scenario YAML files, byte for byte, into
the git-ignored directory `oracle/_ref/` (listed in .gitignore, NOT in .gpurunignore, so it travels to the GPU box
with the snapshot exactly like a compiled oracle would).  Nothing is edited; nothing of it enters the repo's
history; the product package never imports it.  Consumers: `oracle/ref_loader.py` (used by `tests/`,
`bench.py --impl reference` and `bench.py`'s cpu_baseline leg).

    python oracle/make_ref.py            # (re)build oracle/_ref from /root/reference
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("FOT_REFERENCE", "/root/reference")
# what the hot path and its one caller (IntegratedSimulator) import; visualisation / datasets / calibration are not needed
PACKAGES = ("core", "planning", "simulation", "config", "prediction", "pedestrian")


def _copy_tree(src: str, dst: str) -> int:
    n = 0
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(base, src)
        os.makedirs(os.path.join(dst, rel), exist_ok=True)
        for f in files:
            if f.endswith((".pyc", ".pyo")):
                continue
            shutil.copy2(os.path.join(base, f), os.path.join(dst, rel, f))
            n += 1
    return n


def build_ref(verbose: bool = True) -> str | None:
    """Copy the reference into oracle/_ref.  Returns the directory, or None when /root/reference is absent
    (the GPU box: the prebuilt copy that travelled with the snapshot is used as is)."""
    if not os.path.isdir(os.path.join(SOURCE, "src")):
        return DEST if os.path.isdir(os.path.join(DEST, "src")) else None
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(os.path.join(DEST, "src"))
    shutil.copy2(os.path.join(SOURCE, "src", "__init__.py"), os.path.join(DEST, "src", "__init__.py"))
    n = 1
    for pkg in PACKAGES:
        n += _copy_tree(os.path.join(SOURCE, "src", pkg), os.path.join(DEST, "src", pkg))
    os.makedirs(os.path.join(DEST, "scenarios"))
    for f in sorted(os.listdir(os.path.join(SOURCE, "scenarios"))):
        if f.endswith(".yaml"):
            shutil.copy2(os.path.join(SOURCE, "scenarios", f), os.path.join(DEST, "scenarios", f))
            n += 1
    # unmodified: every copied file compares equal to its source
    for pkg in PACKAGES:
        cmp = filecmp.dircmp(os.path.join(SOURCE, "src", pkg), os.path.join(DEST, "src", pkg), ignore=["__pycache__"])
        assert not cmp.diff_files and not cmp.right_only, (pkg, cmp.diff_files, cmp.right_only)
    with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as f:
        f.write(f"byte-for-byte copy of {SOURCE}/src/{{{','.join(PACKAGES)}}} and scenarios/*.yaml made by oracle/make_ref.py; "
                f"{n} files; not part of the repository (git-ignored)\n")
    if verbose:
        print(f"oracle/_ref: {n} files copied from {SOURCE}")
    return DEST


if __name__ == "__main__":
    out = build_ref()
    if out is None:
        print("no reference available (neither /root/reference nor a prebuilt oracle/_ref)")
        sys.exit(1)
