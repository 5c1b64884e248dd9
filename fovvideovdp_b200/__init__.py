"""fovvideovdp_b200 -- a B200-native (sm_100a CUDA) core for the FovVideoVDP per-frame hot path.

Drop-in for the `pyfvvdp` metric API: `fvvdp` (predict / predict_video_source), the display-model plugin
classes and the array video source.  `install()` rebinds the reference package's `fvvdp` class so that the
reference's own command line (`pyfvvdp.run_fvvdp.main`) and examples run on this core unchanged.
"""
from .display_model import (fvvdp_display_geometry, fvvdp_display_photo_absolute, fvvdp_display_photo_eotf, fvvdp_display_photo_gog,
                            fvvdp_display_photometry)
from .fvvdp import fvvdp
from .pupsnr import pu_psnr
from .video_source import fvvdp_video_source, fvvdp_video_source_array, fvvdp_video_source_dm, reshuffle_dims
from .video_source_file import fvvdp_video_source_file, load_image_as_array
from .video_source_yuv import fvvdp_video_source_yuv_file

__all__ = ["fvvdp", "pu_psnr", "fvvdp_display_photometry", "fvvdp_display_photo_eotf", "fvvdp_display_photo_gog", "fvvdp_display_photo_absolute",
           "fvvdp_display_geometry", "fvvdp_video_source", "fvvdp_video_source_dm", "fvvdp_video_source_array", "fvvdp_video_source_yuv_file", "fvvdp_video_source_file", "load_image_as_array", "reshuffle_dims", "install", "uninstall"]


def install(video_sources=True):
    """Make an installed reference package use this core: `pyfvvdp.fvvdp` (and `pyfvvdp.fvvdp.fvvdp`) and `pyfvvdp.pu_psnr`
    become the classes of this package, so the reference's own command line (`pyfvvdp.run_fvvdp.main`, run_fvvdp.py:177-227)
    and examples score on the CUDA core.  With `video_sources` the file source the CLI constructs
    (`pyfvvdp.fvvdp_video_source_file`) is rebound as well: images and raw `.yuv` clips then go through this package's loaders
    and conversion kernels, every other container is handed to the reference's own ffmpeg reader.  Display models and array
    sources of the reference are used as they are.  The original classes stay reachable as `pyfvvdp._reference_classes`."""
    import sys

    import pyfvvdp
    ref_module = sys.modules["pyfvvdp.fvvdp"]  # the package attribute of that name is the class, not the module

    if not hasattr(pyfvvdp, "_reference_classes"):
        pyfvvdp._reference_classes = {"fvvdp": ref_module.fvvdp, "pu_psnr": getattr(pyfvvdp, "pu_psnr", None),
                                      "fvvdp_video_source_file": getattr(pyfvvdp, "fvvdp_video_source_file", None)}
    ref_module.fvvdp = fvvdp
    pyfvvdp.fvvdp = fvvdp
    pyfvvdp.pu_psnr = pu_psnr
    if video_sources:
        pyfvvdp.fvvdp_video_source_file = fvvdp_video_source_file
    return pyfvvdp


def uninstall():
    """Undo install()."""
    import sys

    import pyfvvdp
    ref_module = sys.modules["pyfvvdp.fvvdp"]

    orig = getattr(pyfvvdp, "_reference_classes", None)
    if orig:
        ref_module.fvvdp = pyfvvdp.fvvdp = orig["fvvdp"]
        if orig["pu_psnr"] is not None:
            pyfvvdp.pu_psnr = orig["pu_psnr"]
        if orig["fvvdp_video_source_file"] is not None:
            pyfvvdp.fvvdp_video_source_file = orig["fvvdp_video_source_file"]
        del pyfvvdp._reference_classes
