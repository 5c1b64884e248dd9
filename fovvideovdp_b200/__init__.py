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
           "fvvdp_display_geometry", "fvvdp_video_source", "fvvdp_video_source_dm", "fvvdp_video_source_array", "fvvdp_video_source_yuv_file", "fvvdp_video_source_file", "load_image_as_array", "reshuffle_dims", "install"]


def install():
    """Make an installed reference package use this core: `pyfvvdp.fvvdp` (and `pyfvvdp.fvvdp.fvvdp`) become
    fovvideovdp_b200.fvvdp.  Display models and video sources of the reference are used as they are."""
    import pyfvvdp
    import pyfvvdp.fvvdp as ref_module

    ref_module.fvvdp = fvvdp
    pyfvvdp.fvvdp = fvvdp
    return pyfvvdp
