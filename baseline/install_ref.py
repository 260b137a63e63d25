#!/usr/bin/env python
"""Install the UNMODIFIED reference (facebookresearch/LaViLa at /root/reference) into baseline/_ref/ with pip.

    python baseline/install_ref.py            # no-op when /root/reference is absent (the GPU box) or already installed

baseline/_ref/ is git-ignored (never in history) but NOT gpurun-ignored, so the install travels to the GPU box where
`bench.py --impl reference` (host CPU cores) and the in-run `eager_baseline` (reference modules on the B200 under
autocast, main_pretrain.py:486-530) execute it.  The reference tree has no setup.py / pyproject.toml and is read-only, so
the one-line packaging stub below is written into a scratch COPY under /tmp and pip builds from there
(`--no-index --no-build-isolation --no-deps`: timm / decord / ftfy are not in the wheelhouse; baseline/ref_shim.py stubs
the three symbols the model files import from them).  No reference source is copied into tracked files.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("LAVILA_REFERENCE_ROOT", "/root/reference")

SETUP = """from setuptools import setup, find_namespace_packages
setup(name="lavila", version="0.0.0", packages=find_namespace_packages(include=["lavila", "lavila.*"]),
      package_data={"lavila.models": ["*.gz"]}, zip_safe=False)
"""


def installed():
    return os.path.isfile(os.path.join(TARGET, "lavila", "models", "models.py"))


def install(force=False):
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(REFERENCE, "lavila")):
        return None
    tmp = tempfile.mkdtemp(prefix="lavila_ref_src_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git", "assets", "datasets", "docs"))
        with open(os.path.join(src, "setup.py"), "w") as f:
            f.write(SETUP)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + p.stdout[-2000:] + p.stderr[-2000:])
        # the training / inference drivers are top-level scripts, not package members: keep them beside the package so
        # the baseline can cite the literal train() it mirrors (read-only use)
        for f in ("main_pretrain.py", "main_infer_narrator.py"):
            shutil.copy(os.path.join(REFERENCE, f), os.path.join(TARGET, f))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET


if __name__ == "__main__":
    t = install(force="--force" in sys.argv)
    print("reference installed at", t if t else "(unavailable: %s missing)" % REFERENCE)
