#!/usr/bin/env python
"""Copies the UNMODIFIED reference tree into baseline/_ref/ (git-ignored, NOT gpurun-ignored) so that it travels to the
GPU box, where /root/reference does not exist.

    python baseline/install_ref.py [--src /root/reference]

The reference (liuyuan-pal/NeuRay) has no setup.py / pyproject.toml, so `pip install --target baseline/_ref` does not
apply; it is a plain script tree that is run from its root.  Everything is copied verbatim except `assets/` (11 MB of
README images) and VCS/cache directories.  What uses the copy:
  * tests/test_reference_gpu.py   the reference's own NeuralRay*Renderer under CUDA, patched vs unpatched
  * bench.py --impl reference     the reference's own render() on the box's host cores (cpu_baseline.kind = "reference")
  * bench.py eager_gpu_baseline   the same code under CUDA on the B200
Nothing under neuray_b200/ reads it.
"""
import argparse
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SKIP_TOP = {"assets", ".git", ".github", "__pycache__"}


def install(src="/root/reference", force=False):
    if not os.path.isdir(os.path.join(src, "network")):
        return None
    marker = os.path.join(DST, "network", "renderer.py")
    if not force and os.path.exists(marker) and filecmp.cmp(marker, os.path.join(src, "network", "renderer.py"), shallow=False):
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for name in sorted(os.listdir(src)):
        if name in SKIP_TOP:
            continue
        s, d = os.path.join(src, name), os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    return DST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    out = install(a.src, a.force)
    print(out if out else f"reference tree not found at {a.src}")
    sys.exit(0 if out else 1)
