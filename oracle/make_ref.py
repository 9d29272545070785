#!/usr/bin/env python
"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference's Python sources, so that the GPU box (where
/root/reference does not exist) can execute the REAL reference as the CPU baseline (`bench.py --impl reference`,
`cpu_baseline.kind == "reference"`).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Nothing under contouring_uncertainty_b200/ imports it.  The copy is never committed
(`oracle/_ref/` is in .gitignore, not in .gpurunignore: it travels with the snapshot like the built .so files).  Run by
`__graft_entry__.build()` whenever /root/reference is present; a manifest with the sha256 of every copied file is written
next to the copy so that a stale or edited copy is detectable.

What is copied: the .py files of contour_uncertainty/ and vital/vital/ (the two packages the path imports), nothing else.
The two shape-model .npy files the reference reads are data, not source; the product package ships them already
(contouring_uncertainty_b200/data/) and the runner points the reference's samplers at those.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("CU_REFERENCE_SRC", "/root/reference")
PACKAGES = ("contour_uncertainty", os.path.join("vital", "vital"))


def make(src: str = SRC, dest: str = DEST) -> dict:
    if not os.path.isdir(os.path.join(src, "contour_uncertainty")):
        raise FileNotFoundError(f"reference not found at {src}")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    manifest = {}
    for pkg in PACKAGES:
        for root, _dirs, files in os.walk(os.path.join(src, pkg)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                s = os.path.join(root, f)
                rel = os.path.relpath(s, src)
                d = os.path.join(dest, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                manifest[rel] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    for f in ("LICENSE",):
        if os.path.exists(os.path.join(src, f)):
            shutil.copyfile(os.path.join(src, f), os.path.join(dest, f))
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=0, sort_keys=True)
    return manifest


def verify(dest: str = DEST) -> bool:
    """True when every file of the copy still has the sha256 recorded at copy time."""
    p = os.path.join(dest, "MANIFEST.json")
    if not os.path.exists(p):
        return False
    files = json.load(open(p))["files"]
    for rel, h in files.items():
        f = os.path.join(dest, rel)
        if not os.path.exists(f) or hashlib.sha256(open(f, "rb").read()).hexdigest() != h:
            return False
    return len(files) > 0


if __name__ == "__main__":
    m = make()
    print(f"oracle/_ref: {len(m)} files copied from {SRC}", file=sys.stderr)
