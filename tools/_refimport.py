"""Import the UNMODIFIED reference (pyfvvdp): from git-ignored baseline/_ref/ (installed by tools/vendor_reference.py; it
travels to the GPU box) or, in the build container, from /root/reference.

Only tools/, bench.py (reference arm and `reference_cuda`) and tests/ use this.  No product code imports the reference.

The reference imports `imageio` and `ffmpeg` at module top (video_source_file.py:4,8); they are I/O-only and absent
here, so stub modules are registered first.  `imageio.v2.imread` is backed by OpenCV so that the
reference CLI can load the .png pairs of the drop-in tests."""
import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED = os.path.join(ROOT, "baseline", "_ref")
REFERENCE_ROOT = "/root/reference"


def reference_location():
    if os.path.isdir(os.path.join(VENDORED, "pyfvvdp")):
        return VENDORED
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "pyfvvdp")):
        return REFERENCE_ROOT
    return None


def _imread(path, *args, **kwargs):
    import cv2

    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    return img[:, :, ::-1].copy() if img.ndim == 3 else img


def import_reference():
    loc = reference_location()
    if loc is None:
        raise ImportError("reference not available: run tools/vendor_reference.py in the build container")
    if "imageio" not in sys.modules:
        io = types.ModuleType("imageio")
        v2 = types.ModuleType("imageio.v2")
        v2.imread = _imread
        io.v2 = v2
        io.imread = _imread
        sys.modules["imageio"], sys.modules["imageio.v2"] = io, v2
    if "ffmpeg" not in sys.modules:
        sys.modules["ffmpeg"] = types.ModuleType("ffmpeg")
    if loc not in sys.path:
        sys.path.insert(0, loc)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pyfvvdp  # noqa: F401
    return sys.modules["pyfvvdp"]
