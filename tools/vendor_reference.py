#!/usr/bin/env python
"""Install the UNMODIFIED reference (pyfvvdp 1.2.2) into git-ignored baseline/_ref/ so that it travels to the GPU box:

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>

(`--no-deps`: imageio and ffmpeg-python are I/O-only dependencies that no wheelhouse here carries; they are stubbed at
import time by tools/_refimport.py.  The source tree is read-only, so pip builds from a copy under /tmp.)
bench.py times this package as the reference arm (CPU) and as `reference_cuda` beside our own line; the `-m gpu` tests
use it for drop-in and parity checks.  Nothing of it is tracked in git and no product code imports it."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
SRC = "/root/reference"


def vendor(force=False):
    if os.path.isdir(os.path.join(DEST, "pyfvvdp")) and not force:
        return DEST
    if not os.path.isdir(SRC):
        raise RuntimeError(f"{SRC} not present: the reference can only be vendored in the build container")
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(os.path.dirname(DEST), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns("matlab", "*.mp4", ".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", DEST, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout + r.stderr)
    # the two example images the drop-in tests feed to the reference CLI
    media = os.path.join(DEST, "example_media")
    os.makedirs(media, exist_ok=True)
    for f in ("wavy_facade.png",):
        p = os.path.join(SRC, "example_media", f)
        if os.path.isfile(p):
            shutil.copy2(p, os.path.join(media, f))
    return DEST


if __name__ == "__main__":
    print(vendor(force="-f" in sys.argv))
