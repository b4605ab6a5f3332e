#!/usr/bin/env python3
"""Stage the UNMODIFIED reference modules of the hot path under baseline/_ref/ so they can travel to the GPU box.

    python oracle/build_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (no setup.py / pyproject.toml: not pip-installable, nothing to compile), so its
"build" is a byte-for-byte copy of the files the path needs:

    /root/reference/src/{transformer,ctcModel,utils}/*.py  ->  baseline/_ref/src/...

baseline/_ref/ (the place bench.py's reference arm runs the unmodified reference from) is listed in .gitignore (reference sources never enter this repository's history) and not in
.gpurunignore, so - like the built .so files - it travels with the snapshot to the GPU box, where /root/reference
does not exist.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's reference arms
(`--impl reference`, `--impl reference-gpu`, the `cpu_baseline` leg) may load it, through oracle/ref_loader.py,
and only as the checker / the baseline being measured - never as part of the product path.
"""
import hashlib
import os
import shutil
import sys

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "src")
PACKAGES = ("transformer", "ctcModel", "utils")


def build(verbose=False):
    """Returns the staged directory, or None when the reference is not present (GPU box: use what travelled)."""
    if not os.path.isdir(REF_SRC):
        return DEST if os.path.isdir(DEST) else None
    manifest = []
    for pkg in PACKAGES:
        src_dir, dst_dir = os.path.join(REF_SRC, pkg), os.path.join(DEST, pkg)
        os.makedirs(dst_dir, exist_ok=True)
        for name in sorted(os.listdir(src_dir)):
            if not name.endswith(".py"):
                continue
            src, dst = os.path.join(src_dir, name), os.path.join(dst_dir, name)
            shutil.copyfile(src, dst)
            with open(dst, "rb") as f:
                manifest.append("%s  %s/%s" % (hashlib.sha256(f.read()).hexdigest(), pkg, name))
    with open(os.path.join(os.path.dirname(DEST), "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print("staged %d reference files under %s" % (len(manifest), DEST))
    return DEST


if __name__ == "__main__":
    path = build(verbose=True)
    if path is None:
        print("reference not present and nothing staged", file=sys.stderr)
        sys.exit(1)
