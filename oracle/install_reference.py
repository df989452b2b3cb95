#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE.  Install the UNMODIFIED reference package into baseline/_ref/ so that bench.py's
CPU arm (`--impl reference`) can time the reference's own `step()` (SPaRC_Gym/SPaRC_Gym.py:1111-1238) on the GPU
box, where /root/reference does not exist.  baseline/_ref is git-ignored (never in history) but NOT gpurun-ignored.

    python oracle/install_reference.py [--src /root/reference]

1. the contract's recipe: `pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
   --target baseline/_ref <copy of the source tree>`.  In this image it fails: the reference's build backend
   (hatchling, pyproject.toml:1-3) is neither installed nor in the wheelhouse.
2. fallback = what that wheel would contain (`[tool.hatch.build.targets.wheel] packages = ["SPaRC_Gym"]`,
   pyproject.toml:31-32): the package directory, byte for byte, plus a MANIFEST with the sha256 of every file.
Nothing is modified; the harness (oracle/refharness.py) supplies stub `gymnasium` / `pygame` modules and the rows.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(src="/root/reference", dest=DEST, verbose=True):
    """Returns a dict describing what was done (also written to <dest>/INSTALL.json)."""
    if not os.path.isfile(os.path.join(src, "SPaRC_Gym", "SPaRC_Gym.py")):
        raise FileNotFoundError(f"no reference tree under {src}")
    shutil.rmtree(dest, ignore_errors=True)
    os.makedirs(dest, exist_ok=True)
    how, pip_err = "pip", ""
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "src")
        shutil.copytree(src, work)                       # the source tree is read-only: build from a copy
        res = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                              "--find-links", "/opt/wheelhouse", "--target", dest, work], capture_output=True, text=True)
        if res.returncode != 0 or not os.path.isfile(os.path.join(dest, "SPaRC_Gym", "SPaRC_Gym.py")):
            how = "package copy (pip failed)"
            pip_err = (res.stderr or res.stdout).strip().splitlines()[-1:] or [""]
            pip_err = pip_err[0][:200]
            shutil.rmtree(dest, ignore_errors=True)
            os.makedirs(dest, exist_ok=True)
            shutil.copytree(os.path.join(src, "SPaRC_Gym"), os.path.join(dest, "SPaRC_Gym"),
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    files = {}
    for d, _, names in os.walk(os.path.join(dest, "SPaRC_Gym")):
        for n in names:
            if n.endswith(".pyc"):
                continue
            p = os.path.join(d, n)
            files[os.path.relpath(p, dest)] = _sha(p)
    # unmodified: every installed file equals its source
    for rel, digest in files.items():
        s = os.path.join(src, rel)
        if os.path.isfile(s) and _sha(s) != digest:
            raise RuntimeError(f"{rel} differs from the reference source")
    info = {"source": src, "how": how, "pip_error": pip_err, "files": files}
    with open(os.path.join(dest, "INSTALL.json"), "w") as f:
        json.dump(info, f, indent=1, sort_keys=True)
    if verbose:
        print(f"reference installed into {dest}: {how}; {len(files)} files" + (f" [{pip_err}]" if pip_err else ""))
    return info


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    install(a.src)
