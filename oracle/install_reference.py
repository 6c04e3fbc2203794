#!/usr/bin/env python
"""Installs the UNMODIFIED reference (TeemoCaption/FDDM-asr, a plain source tree without packaging metadata) into
baseline/_ref so that bench.py's baseline arms (`cpu_baseline`, `--impl reference`, `eager_b200`) can run the
reference's own classes on the GPU box, where /root/reference does not exist.

    python oracle/install_reference.py [--ref /root/reference]

Recipe (the base contract's): the tree is read-only and has no setup.py / pyproject.toml, so it is copied to a
temporary directory, a six-line packaging shim (NOT part of the reference) is added there, and
    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>
installs it.  baseline/_ref is git-ignored (no reference source enters the history) but not gpurun-ignored.
The installed .py files are byte-identical to the reference's; this script verifies that."""
import argparse
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = '''# Packaging shim written by the build (NOT part of the reference).
from setuptools import setup, find_namespace_packages
setup(name="fddm-asr-reference", version="0.0.0",
      packages=find_namespace_packages(include=["fddm", "fddm.*", "models", "losses", "sampler"]),
      py_modules=["train", "inference"])
'''
CHECK = ["fddm/sched/diffusion_scheduler.py", "losses/fddm_losses.py", "sampler/jumpy_sampler.py", "train.py",
         "models/evaluate.py", "models/projection.py"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("FDDM_REF", "/root/reference"))
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    dst = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(args.ref):
        print(f"reference tree {args.ref} not present: nothing to install")
        return 0
    if os.path.exists(os.path.join(dst, "fddm", "sched", "diffusion_scheduler.py")) and not args.force:
        print(f"{dst} already holds the reference")
    else:
        tmp = tempfile.mkdtemp(prefix="fddm_ref_")
        try:
            src = os.path.join(tmp, "ref")
            shutil.copytree(args.ref, src, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".git"))
            with open(os.path.join(src, "setup.py"), "w") as f:
                f.write(SHIM)
            shutil.rmtree(dst, ignore_errors=True)
            os.makedirs(dst, exist_ok=True)
            subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", dst, src], check=True, cwd=src)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    bad = [f for f in CHECK if not filecmp.cmp(os.path.join(args.ref, f), os.path.join(dst, f), shallow=False)]
    if bad:
        raise SystemExit(f"installed files differ from the reference: {bad}")
    print(f"reference installed in {dst}; {len(CHECK)} files verified byte-identical")
    return 0


if __name__ == "__main__":
    sys.exit(main())
