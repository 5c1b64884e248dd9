"""File-based frame providers.  `fvvdp_video_source_file` keeps the reference's constructor (pyfvvdp/video_source_file.py:
413-443): it recognises what the files are and wraps the matching provider --

  * images (.png .jpg .jpeg .bmp .gif .ppm .tiff)  -> fvvdp_video_source_array, dim_order "HWC" (16-bit PNGs stay 16 bit)
  * raw planar .yuv clips                            -> fvvdp_video_source_yuv_file (one CUDA conversion kernel per frame)
  * anything else is a container the reference decodes through an ffmpeg pipe (video_source_file.py:59-164); that reader
    stays with the reference (I/O plumbing is out of scope), so such files raise here.
"""
import logging
import os

import numpy as np

from .video_source import fvvdp_video_source, fvvdp_video_source_array
from .video_source_yuv import fvvdp_video_source_yuv_file

IMAGE_EXTENSIONS = [".png", ".jpg", ".gif", ".bmp", ".jpeg", ".ppm", ".tiff", ".tif"]


def load_image_as_array(imgfile):
    """(H,W,C) uint8 / uint16 array, RGB order, alpha dropped, (H,W,1) for grey images (video_source_file.py:29-54)."""
    img = None
    try:
        import cv2
        img = cv2.imread(imgfile, cv2.IMREAD_UNCHANGED)
        if img is not None and img.ndim == 3:
            img = img[:, :, [2, 1, 0] + list(range(3, img.shape[2]))]  # BGR(A) -> RGB(A)
    except ImportError:
        pass
    if img is None:
        ext = os.path.splitext(imgfile)[1].lower()
        if ext in (".exr", ".hdr", ".dds"):
            raise RuntimeError(f'"{imgfile}": {ext} images need an HDR-capable reader (OpenCV with OpenEXR, or imageio); none is available here')
        from PIL import Image
        im = Image.open(imgfile)
        if im.mode in ("P", "PA", "LA", "1"):  # palette indices / grey+alpha are not pixel values
            im = im.convert("RGB" if im.mode in ("P", "PA") else "L")
        elif im.mode.startswith("I;16") or im.mode == "I":
            pass  # 16-bit grey survives np.array()
        img = np.array(im)
        if im.mode == "I":
            img = img.astype(np.uint16)
    if img.ndim == 3 and img.shape[2] > 3:
        logging.warning(f"Input image {imgfile} has more than 3 channels (alpha?). Ignoring the extra channels.")
        img = img[:, :, :3]
    if img.ndim == 2:
        img = img[:, :, np.newaxis]
    return np.ascontiguousarray(img)


class fvvdp_video_source_file(fvvdp_video_source):
    def __init__(self, test_fname, reference_fname, display_photometry="sdr_4k_30", color_space_name="auto", frames=-1,
                 full_screen_resize=None, resize_resolution=None, preload=False, ffmpeg_cc=False, verbose=False):
        assert os.path.isfile(test_fname), f'File does not exists: "{test_fname}"'
        assert os.path.isfile(reference_fname), f'File does not exists: "{reference_fname}"'
        ext_t, ext_r = os.path.splitext(test_fname)[1].lower(), os.path.splitext(reference_fname)[1].lower()
        if ext_t in IMAGE_EXTENSIONS or ext_t in (".exr", ".hdr", ".dds"):
            assert ext_r in IMAGE_EXTENSIONS, "Test is an image, but reference is a video"
            if color_space_name == "auto":
                color_space_name = "sRGB"
            if full_screen_resize is not None:
                logging.error("full-screen-resize not implemented for images.")
            self.vs = fvvdp_video_source_array(load_image_as_array(test_fname), load_image_as_array(reference_fname), 0, dim_order="HWC",
                                               display_photometry=display_photometry, color_space_name=color_space_name)
        else:
            assert ext_r not in IMAGE_EXTENSIONS, "Test is a video, but reference is an image"
            if ext_t == ".yuv" and ext_r == ".yuv":
                self.vs = fvvdp_video_source_yuv_file(test_fname, reference_fname, display_photometry=display_photometry,
                                                      color_space_name=color_space_name, frames=frames, full_screen_resize=full_screen_resize,
                                                      resize_resolution=resize_resolution, verbose=verbose)
            else:
                # any other container: the reference's own ffmpeg reader, when that package is installed (plumbing that stays
                # with the reference); it hands luminance frames to the metric through get_*_frame()
                ref_cls = None
                try:
                    import pyfvvdp
                    ref_cls = getattr(pyfvvdp, "_reference_classes", {}).get("fvvdp_video_source_file") or pyfvvdp.fvvdp_video_source_file
                except ImportError:
                    pass
                if ref_cls is None or ref_cls is fvvdp_video_source_file:
                    raise RuntimeError(f'"{test_fname}" needs the reference\'s ffmpeg reader (pyfvvdp.fvvdp_video_source_file); pass that '
                                       "source to predict_video_source() or convert the clip to raw .yuv")
                self.vs = ref_cls(test_fname, reference_fname, display_photometry=display_photometry, color_space_name=color_space_name,
                                  frames=frames, full_screen_resize=full_screen_resize, resize_resolution=resize_resolution, preload=preload,
                                  ffmpeg_cc=ffmpeg_cc, verbose=verbose)

    def get_video_size(self):
        return self.vs.get_video_size()

    def get_frames_per_second(self):
        return self.vs.get_frames_per_second()

    def get_test_frame(self, frame, device):
        return self.vs.get_test_frame(frame, device)

    def get_reference_frame(self, frame, device):
        return self.vs.get_reference_frame(frame, device)
