"""Frame providers.  Interface and names follow the reference's pyfvvdp/video_source.py:

  fvvdp_video_source        abstract provider (:14-36): get_video_size() -> (H, W, N), get_frames_per_second(),
                            get_test_frame(i, device) / get_reference_frame(i, device) -> float32 luminance
                            (1,1,1,H,W) in cd/m^2 on `device`
  reshuffle_dims            (:43-69)
  fvvdp_video_source_dm     provider with a photometric display model + RGB->Y weights (:75-93)
  fvvdp_video_source_array  numpy / torch arrays in any dimension order (:104-208)

The metric's CUDA front end reads the raw arrays of an fvvdp_video_source_array directly (uint8 / uint16 /
float32, 1 or 3 channels, any strides) and applies EOTF + RGB->Y in-kernel; get_*_frame() below is the
torch statement of the same arithmetic, used when some other consumer asks for luminance frames.
"""
import numpy as np
import torch

from . import config
from .display_model import fvvdp_display_photometry


class fvvdp_video_source:
    def get_video_size(self):
        raise NotImplementedError

    def get_frames_per_second(self):
        raise NotImplementedError

    def get_test_frame(self, frame, device):
        raise NotImplementedError

    def get_reference_frame(self, frame, device):
        raise NotImplementedError


def reshuffle_dims(T, in_dims, out_dims):
    """Permute `T` from the dimension order `in_dims` (e.g. "HWC") to `out_dims` (e.g. "BCFHW"); dimensions
    absent from in_dims become singletons.  Returns a view whenever torch can."""
    in_dims, out_dims = in_dims.upper(), out_dims.upper()
    for ch in in_dims:
        if ch not in out_dims:
            raise RuntimeError('Dimension "{}" missing in the target dimensions: "{}"'.format(ch, out_dims))
    present = [ch for ch in out_dims if ch in in_dims]
    Tp = T.permute([in_dims.index(ch) for ch in present])
    shape = [Tp.shape[present.index(ch)] if ch in present else 1 for ch in out_dims]
    return Tp.reshape(shape)


class fvvdp_video_source_dm(fvvdp_video_source):
    def __init__(self, display_photometry="sdr_4k_30", color_space_name="sRGB"):
        self.color_to_luminance = config.rgb2y(color_space_name)
        if isinstance(display_photometry, str):
            self.dm_photometry = fvvdp_display_photometry.load(display_photometry)
        elif hasattr(display_photometry, "forward"):
            self.dm_photometry = display_photometry
        else:
            raise RuntimeError("display_model must be a string or fvvdp_display_photometry subclass")


def _as_tensor(a):
    if isinstance(a, np.ndarray):
        if a.dtype == np.uint16:
            a = a.view(np.int16)  # torch has no uint16 arithmetic: same bits, unpacked with & 0xFFFF downstream
        return torch.from_numpy(a if a.flags.writeable else a.copy())
    return a


class fvvdp_video_source_array(fvvdp_video_source_dm):
    """first_frame / total_frames (extensions): the arrays hold frames [first_frame, first_frame + F) of a clip of
    total_frames frames -- what one rank keeps when a clip is sharded over several GPUs (its frame block plus the
    temporal halo before it).  Frame indices passed to get_*_frame() are clip indices."""

    def __init__(self, test_video, reference_video, fps, dim_order="BCFHW", display_photometry="sdr_4k_30", color_space_name="sRGB",
                 first_frame=0, total_frames=None):
        super().__init__(display_photometry=display_photometry, color_space_name=color_space_name)
        if tuple(test_video.shape) != tuple(reference_video.shape):
            raise RuntimeError("Test and reference image/video tensors must be exactly the same shape")
        if len(dim_order) != len(test_video.shape):
            raise RuntimeError('Input tensor much have exactly as many dimensions as there are characters in the "dims" parameter')
        test_video = reshuffle_dims(_as_tensor(test_video), dim_order, "BCFHW")
        reference_video = reshuffle_dims(_as_tensor(reference_video), dim_order, "BCFHW")
        B, C, F, H, W = test_video.shape
        if fps == 0 and F > 1:
            raise RuntimeError("When passing video sequences, you must set 'frames_per_second' parameter")
        if C != 3 and C != 1:
            raise RuntimeError("The content must have either 1 or 3 colour channels.")
        self.fps = fps
        self.is_video = fps > 0
        self.is_color = C == 3
        self.test_video = test_video
        self.reference_video = reference_video
        self.first_frame = int(first_frame)
        self.total_frames = int(total_frames) if total_frames is not None else F + self.first_frame
        if self.total_frames < self.first_frame + F:
            raise RuntimeError("total_frames is smaller than first_frame + the number of frames held")
        if self.total_frames > 1 and fps == 0:
            raise RuntimeError("When passing video sequences, you must set 'frames_per_second' parameter")

    def local_index(self, frame):
        k = frame - self.first_frame
        if k < 0 or k >= self.test_video.shape[2]:
            raise RuntimeError(f"frame {frame} is not held by this process (it holds [{self.first_frame}, "
                               f"{self.first_frame + self.test_video.shape[2]}))")
        return k

    def get_frames_per_second(self):
        return self.fps

    def get_video_size(self):
        sh = self.test_video.shape
        return (sh[3], sh[4], self.total_frames)

    def get_test_frame(self, frame, device=torch.device("cpu")):
        return self._get_frame(self.test_video, frame, device)

    def get_reference_frame(self, frame, device=torch.device("cpu")):
        return self._get_frame(self.reference_video, frame, device)

    def _get_frame(self, from_array, frame, device):
        k = self.local_index(frame)
        V = from_array[:, :, k:k + 1].to(device)
        if V.dtype == torch.float32:
            pass
        elif V.dtype == torch.int16:
            V = (V.to(torch.int32) & 0xFFFF).to(torch.float32) / 65535
        elif V.dtype == torch.uint8:
            V = V.to(torch.float32) / 255
        else:
            raise RuntimeError("Only uint8, uint16 and float32 is currently supported")
        L = self.dm_photometry.forward(V)
        if self.is_color:
            w = self.color_to_luminance
            L = L[:, 0:1] * w[0] + L[:, 1:2] * w[1] + L[:, 2:3] * w[2]
        return L


def is_array_source(vs):
    """An array-backed provider of ours or of the reference package (duck-typed on the attributes that
    pyfvvdp/video_source.py:151-155 sets)."""
    return all(hasattr(vs, a) for a in ("test_video", "reference_video", "dm_photometry", "color_to_luminance", "is_color")) and \
        torch.is_tensor(getattr(vs, "test_video")) and type(vs).__name__ == "fvvdp_video_source_array"
