"""Import the UNMODIFIED reference (pyfvvdp) from /root/reference in the build container.

Only tools/ scripts use this (to generate golden vectors and data files). Nothing under
tests/ -m gpu, bench.py or the product package may import it: /root/reference does not exist
on the GPU box.

The reference imports `imageio` and `ffmpeg` at module top (video_source_file.py:4,8); they are
I/O-only and absent here, so empty stub modules are registered first.
"""
import sys
import types
import warnings

REFERENCE_ROOT = "/root/reference"


def import_reference():
    for name in ("imageio", "imageio.v2", "ffmpeg"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pyfvvdp  # noqa: F401
    return sys.modules["pyfvvdp"]
