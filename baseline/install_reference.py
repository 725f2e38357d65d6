"""Copy the unmodified reference package into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it ships to the GPU box).

The reference has no packaging metadata (no setup.py / pyproject.toml), so `pip install /root/reference` does not apply;
its `letf/` package is pure Python and importable as a directory.  Nothing is edited.

usage: python baseline/install_reference.py [--src /root/reference]
"""
from __future__ import annotations

import argparse
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def install(src: str = "/root/reference") -> bool:
    pkg = os.path.join(src, "letf")
    if not os.path.isdir(pkg):
        return False
    os.makedirs(DST, exist_ok=True)
    out = os.path.join(DST, "letf")
    if os.path.isdir(out):
        shutil.rmtree(out)
    shutil.copytree(pkg, out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "SOURCE"), "w") as f:
        f.write(f"copied unmodified from {pkg}\n")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    print("installed" if install(a.src) else "reference not found", DST)
