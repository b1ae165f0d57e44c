#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE ONLY.  Recipe for ``oracle/_ref/``: a verbatim, UNMODIFIED copy of the reference's own
Python modules on the reconstruction path, taken from the read-only mount ``/root/reference`` where they lie.

The reference is pure Python (no build system, nothing to compile, nothing pip-installable), so "building" it means
copying the files it imports on the path: ``lib/*.py`` and ``lib/model/*.py`` (``lib/renderer`` and ``lib/data`` are
not on the path and need OpenGL / trimesh).  ``oracle/_ref/`` is git-ignored -- reference SOURCES never enter this
repository's history -- but NOT gpurun-ignored, so the copy travels to the GPU box like the built ``.so`` files and
``bench.py --impl reference`` can time the reference's own ``lib.mesh_util.reconstruction`` on the box's host cores
(``cpu_baseline.kind = "reference"``).  ``oracle/ref_manifest.json`` (committed: paths + sha256, no source) lets
anybody check that what ran is byte-identical to the mount.

    python oracle/make_ref.py            # (re)creates oracle/_ref from /root/reference
    python oracle/make_ref.py --check    # verifies oracle/_ref against oracle/ref_manifest.json
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SURS_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(HERE, "ref_manifest.json")
SUBDIRS = ["lib", os.path.join("lib", "model")]          # flat copies of *.py in these directories


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _files(root):
    out = []
    for sub in SUBDIRS:
        d = os.path.join(root, sub)
        if os.path.isdir(d):
            out += [os.path.join(sub, fn) for fn in sorted(os.listdir(d)) if fn.endswith(".py")]
    return out


def available():
    return os.path.isdir(os.path.join(DST, "lib", "model"))


def make(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "lib")):
        raise RuntimeError("reference mount %s not present" % SRC)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for rel in _files(SRC):
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
        manifest[rel] = _sha(os.path.join(DST, rel))
    for extra in ("LICENSE.txt",):
        if os.path.exists(os.path.join(SRC, extra)):
            shutil.copyfile(os.path.join(SRC, extra), os.path.join(DST, extra))
    with open(MANIFEST, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: %d reference files copied unmodified from %s" % (len(manifest), SRC))
    return manifest


def check():
    """True when every file of oracle/_ref has the sha256 recorded when it was copied from the mount."""
    if not available() or not os.path.exists(MANIFEST):
        return False
    with open(MANIFEST) as f:
        manifest = json.load(f)
    return sorted(manifest) == sorted(_files(DST)) and all(_sha(os.path.join(DST, rel)) == h for rel, h in manifest.items())


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = check()
        print("oracle/_ref matches oracle/ref_manifest.json:", ok)
        sys.exit(0 if ok else 1)
    make()
