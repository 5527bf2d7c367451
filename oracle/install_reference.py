"""Install the UNMODIFIED reference (cinemere/evacuation) into baseline/_ref so that it travels to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (the CPU arm of bench.py and the golden generator); nothing under
evacuation_b200/ imports it.  The reference is pure Python without packaging metadata (`pip install /root/reference` fails:
"does not appear to be a Python project"), so -- as the bench contract allows -- the install is made from a copy under
/tmp to which a two-line setup.py (packaging metadata only, `packages=find_packages(include=["src", "src.*"])`) is
added; the reference's own files are copied byte for byte and nothing of it enters the git history (baseline/_ref is
git-ignored, not gpurun-ignored).  Run in the authoring container, where /root/reference is mounted:

    python oracle/install_reference.py            # -> baseline/_ref/src/env/...
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
SOURCE = os.environ.get("EVAC_REFERENCE_SOURCE", "/root/reference")

_SETUP_PY = ('from setuptools import setup, find_packages\n'
             'setup(name="evacuation-reference", version="0", packages=find_packages(include=["src", "src.*"]))\n')


def installed(target: str = TARGET) -> bool:
    return os.path.isfile(os.path.join(target, "src", "env", "env", "area.py"))


def install(force: bool = False) -> str:
    """Returns a one-line outcome (also recorded in DESIGN.md)."""
    if installed() and not force:
        return f"already installed: {TARGET}"
    if not os.path.isdir(os.path.join(SOURCE, "src", "env")):
        return f"reference sources not found under {SOURCE}: nothing installed (the CPU arm falls back to the oracle port)"
    tmp = tempfile.mkdtemp(prefix="evac_ref_src_")
    try:
        shutil.copytree(os.path.join(SOURCE, "src"), os.path.join(tmp, "src"), ignore=shutil.ignore_patterns("__pycache__"))
        with open(os.path.join(tmp, "setup.py"), "w") as f:
            f.write(_SETUP_PY)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(os.path.dirname(TARGET), exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", TARGET, tmp]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not installed():
            return f"pip install failed (rc {res.returncode}): {res.stderr.strip().splitlines()[-1] if res.stderr.strip() else res.stdout[-200:]}"
        return f"installed the unmodified reference into {TARGET} (pip --no-deps --target, from a /tmp copy + 2-line setup.py)"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
