#!/usr/bin/env python
"""Install the UNMODIFIED reference (CrawfordGroup/pycc) under ``baseline/_ref`` so that it travels to the GPU box.

    python baseline/install_ref.py            # run in the build container (needs /root/reference)

``baseline/_ref`` is git-ignored (the reference is not product source and never enters this repository's history)
but it is not gpurun-ignored, so the snapshot sent to a B200 box carries it.  ``bench.py --impl reference`` and the
``cpu_baseline`` leg import the reference from there through ``baseline/refload.py``.

Two routes, tried in order:

1. the contract's offline install, ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref``
   from a copy of the source tree under /tmp (the tree itself is read-only).  Works in this image (the package
   installs as pycc 0.0.0: no git metadata for its version scheme).
2. fallback, should the build backend refuse: what that install produces for a pure-Python package -- the ``pycc/``
   package directory copied verbatim (``*.py`` only, tests and data left out).

Either way a ``PROVENANCE`` file records the route and a hash of the four hot-path files, which are asserted to be
byte-identical to the source tree.

No file is edited on the way.  The reference's third-party imports that are absent here (psi4, opt_einsum,
qcelemental) are shimmed at load time by ``refload.load_reference``, not by patching sources.
"""
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def _pip(src):
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(src, work, ignore=shutil.ignore_patterns(".git", "docs"))
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DEST, work]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return r.returncode == 0 and os.path.exists(os.path.join(DEST, "pycc", "ccwfn.py")), r.stdout[-400:]


def _copy(src):
    dst = os.path.join(DEST, "pycc")
    os.makedirs(dst, exist_ok=True)
    n = 0
    for root, dirs, files in os.walk(os.path.join(src, "pycc")):
        dirs[:] = [d for d in dirs if d not in ("tests", "data", "__pycache__")]
        rel = os.path.relpath(root, os.path.join(src, "pycc"))
        os.makedirs(os.path.join(dst, rel), exist_ok=True)
        for f in files:
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel, f))
                n += 1
    return n


def install(src="/root/reference", quiet=False):
    """Returns a one-line description of what was done ('' if the reference tree is absent)."""
    if not os.path.isdir(os.path.join(src, "pycc")):
        return ""
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST, exist_ok=True)
    ok, log = _pip(src)
    how = "pip install --no-index --no-build-isolation --no-deps --target baseline/_ref"
    if not ok:
        shutil.rmtree(DEST, ignore_errors=True)
        os.makedirs(DEST, exist_ok=True)
        n = _copy(src)
        how = ("verbatim copy of %d pycc/*.py files (pip install failed: %s)"
               % (n, " ".join(log.split())[-160:] or "no output"))
    sha = hashlib.sha256()
    for name in ("ccwfn.py", "cctriples.py", "utils.py", "device.py"):
        with open(os.path.join(DEST, "pycc", name), "rb") as f, open(os.path.join(src, "pycc", name), "rb") as g:
            a, b = f.read(), g.read()
            assert a == b, "installed %s differs from the reference source" % name
            sha.update(a)
    with open(os.path.join(DEST, "PROVENANCE"), "w") as f:
        f.write("source: %s\nhow: %s\nsha256(ccwfn,cctriples,utils,device): %s\n" % (src, how, sha.hexdigest()))
    if not quiet:
        print("baseline/_ref: " + how)
    return how


if __name__ == "__main__":
    if not install():
        raise SystemExit("no reference tree at /root/reference")
