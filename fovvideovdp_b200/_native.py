"""ctypes binding of libfvvdp_b200.so (C ABI: include/fvvdp_b200.h).  No torch types cross this boundary:
device buffers are passed as integer addresses (tensor.data_ptr()).  There is NO CPU fallback: if the
library is missing and cannot be built, or no CUDA device exists, the caller gets a RuntimeError."""
import ctypes as C
import os
import threading

from . import build as _build

ABI_VERSION = 1
MAX_LEVELS = 16
MAX_FILTER_LEN = 32
MAX_BLOCK_FRAMES = 96
MAX_SLOTS = MAX_BLOCK_FRAMES + MAX_FILTER_LEN

EOTF_CODES = {"none": 0, "sRGB": 1, "gamma": 2, "PQ": 3, "linear": 4, "absolute": 5}
DTYPE_F32, DTYPE_U8, DTYPE_U16 = 0, 1, 2
TAP_R, TAP_GAUSS, TAP_CONTRAST, TAP_LBKG, TAP_S, TAP_D, TAP_DMAP_BAND = range(7)

EXPORTS = ["fvvdp_b200_create", "fvvdp_b200_score_block", "fvvdp_b200_heatmap", "fvvdp_b200_read_tap",
           "fvvdp_b200_level_size", "fvvdp_b200_launch_count", "fvvdp_b200_traffic_model", "fvvdp_b200_destroy",
           "fvvdp_b200_last_error", "fvvdp_b200_abi_version", "fvvdp_b200_pool_jod",
           "fvvdp_b200_profile", "fvvdp_b200_profile_read", "fvvdp_b200_heatmap_visualize", "fvvdp_b200_set_foveation_maps",
           "fvvdp_b200_yuv_to_luminance", "fvvdp_b200_pu_sq_err", "fvvdp_b200_pu_sq_err_frames", "fvvdp_b200_score_block_yuv"]
COLORMAPS = {"threshold": 0, "supra-threshold": 1}
PROFILE_CLASSES = MAX_LEVELS + 2


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("n_levels", C.c_int32),
        ("band_freq", C.c_float * MAX_LEVELS),
        ("temp_ch", C.c_int32),
        ("filter_len", C.c_int32),
        ("filt", (C.c_float * MAX_FILTER_LEN) * 2),
        ("eotf", C.c_int32),
        ("Y_peak", C.c_float), ("Y_black", C.c_float), ("gamma", C.c_float), ("L_min", C.c_float), ("L_max", C.c_float),
        ("rgb2y", C.c_float * 3),
        ("in_dtype", C.c_int32),
        ("in_channels", C.c_int32),
        ("csf_rho_log", C.c_void_p), ("csf_Y_log", C.c_void_p), ("csf_ecc_sqrt", C.c_void_p), ("csf_S_log", C.c_void_p),
        ("csf_rho_range", C.c_float * 2), ("csf_Y_range", C.c_float * 2), ("csf_ecc_range", C.c_float * 2),
        ("mask_p", C.c_float), ("mask_q", C.c_float * 2), ("mask_c_mul", C.c_float),
        ("sens_mul", C.c_float),
        ("beta", C.c_float),
        ("w_transient", C.c_float),
        ("foveated", C.c_int32),
        ("display_size_m", C.c_float * 2), ("distance_m", C.c_float), ("ppd_centre", C.c_float),
        ("want_dmap", C.c_int32),
        ("want_taps", C.c_int32),
        ("max_block_frames", C.c_int32),
    ]


class YuvDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bit_depth", C.c_int32), ("chroma_420", C.c_int32),
                ("ycbcr2rgb", C.c_float * 9), ("eotf", C.c_int32),
                ("Y_peak", C.c_float), ("Y_black", C.c_float), ("gamma", C.c_float), ("L_min", C.c_float), ("L_max", C.c_float),
                ("rgb2y", C.c_float * 3), ("resize", C.c_int32), ("out_width", C.c_int32), ("out_height", C.c_int32)]


RESIZE_CODES = {None: 0, "nearest": 1, "bilinear": 2, "bicubic": 3, "area": 4}  # fvvdp_b200_resize


class PuParams(C.Structure):
    _fields_ = [("p", C.c_float * 7), ("L_min", C.c_float), ("L_max", C.c_float)]


class PoolParams(C.Structure):
    _fields_ = [("beta_sch", C.c_float), ("beta_tch", C.c_float), ("beta_t", C.c_float), ("w_transient", C.c_float),
                ("jod_a", C.c_float), ("log_jod_exp", C.c_float)]


_lib = None


def library_path():
    return _build.LIB


_load_lock = threading.Lock()


def load_library():
    """dlopen libfvvdp_b200.so (building it first if the sources are newer).  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _load_lock:  # worker threads of run_fvvdp.score_pairs may arrive together: one of them builds / loads
        return _load_library_locked()


def _load_library_locked():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.is_stale():
        path = _build.build_native()
    if not os.path.isfile(path):
        raise RuntimeError(f"{path} is missing and could not be built; fovvideovdp_b200 has no CPU fallback")
    lib = C.CDLL(path)
    lib.fvvdp_b200_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]
    lib.fvvdp_b200_create.restype = C.c_int
    lib.fvvdp_b200_score_block.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
                                           C.POINTER(C.c_float), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.fvvdp_b200_score_block.restype = C.c_int
    lib.fvvdp_b200_heatmap.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    lib.fvvdp_b200_heatmap.restype = C.c_int
    lib.fvvdp_b200_heatmap_visualize.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    lib.fvvdp_b200_heatmap_visualize.restype = C.c_int
    lib.fvvdp_b200_set_foveation_maps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.fvvdp_b200_set_foveation_maps.restype = C.c_int
    lib.fvvdp_b200_yuv_to_luminance.argtypes = [C.POINTER(YuvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.fvvdp_b200_yuv_to_luminance.restype = C.c_int
    lib.fvvdp_b200_pu_sq_err.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(PuParams), C.c_void_p, C.c_int, C.c_void_p]
    lib.fvvdp_b200_pu_sq_err.restype = C.c_int
    lib.fvvdp_b200_read_tap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    lib.fvvdp_b200_read_tap.restype = C.c_int64
    lib.fvvdp_b200_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.fvvdp_b200_level_size.restype = C.c_int
    lib.fvvdp_b200_launch_count.argtypes = [C.c_void_p]
    lib.fvvdp_b200_launch_count.restype = C.c_int64
    lib.fvvdp_b200_traffic_model.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.fvvdp_b200_traffic_model.restype = C.c_int
    lib.fvvdp_b200_destroy.argtypes = [C.c_void_p]
    lib.fvvdp_b200_destroy.restype = C.c_int
    lib.fvvdp_b200_last_error.argtypes = [C.c_void_p]
    lib.fvvdp_b200_last_error.restype = C.c_char_p
    lib.fvvdp_b200_abi_version.argtypes = []
    lib.fvvdp_b200_abi_version.restype = C.c_int
    lib.fvvdp_b200_profile.argtypes = [C.c_void_p, C.c_int]
    lib.fvvdp_b200_profile.restype = C.c_int
    lib.fvvdp_b200_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    lib.fvvdp_b200_profile_read.restype = C.c_int
    lib.fvvdp_b200_pool_jod.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.POINTER(PoolParams), C.c_int, C.c_void_p, C.c_void_p]
    lib.fvvdp_b200_pool_jod.restype = C.c_int
    if lib.fvvdp_b200_abi_version() != ABI_VERSION:
        raise RuntimeError("libfvvdp_b200.so ABI version mismatch; rebuild with python -m fovvideovdp_b200.build")
    _lib = lib
    return lib


def last_error(handle=None):
    return load_library().fvvdp_b200_last_error(handle).decode("utf-8", "replace")


class Context:
    """RAII wrapper of fvvdp_b200_ctx."""

    def __init__(self, cfg: Config, device_index: int, keepalive=()):
        self._lib = load_library()
        self._keep = keepalive
        h = C.c_void_p()
        rc = self._lib.fvvdp_b200_create(C.byref(cfg), int(device_index), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"fvvdp_b200_create failed ({rc}): {last_error(None)}")
        self.handle = h
        self.cfg = cfg

    def close(self):
        if getattr(self, "handle", None):
            self._lib.fvvdp_b200_destroy(self.handle)
            self.handle = None

    __del__ = close

    def _check(self, rc, what):
        if rc < 0:
            raise RuntimeError(f"{what} failed ({rc}): {last_error(self.handle)}")
        return rc

    def score_block(self, test_ptrs, ref_ptrs, strides, n_frames, fixation_xy, q_ptr, q_stride, q_col0, flags_ptr, stream):
        n = len(test_ptrs)
        assert n == len(ref_ptrs)
        tp = (C.c_void_p * n)(*test_ptrs)
        rp = (C.c_void_p * n)(*ref_ptrs)
        st = (C.c_int64 * 3)(*[int(s) for s in strides])
        fx = None
        if fixation_xy is not None:
            flat = [float(v) for xy in fixation_xy for v in xy]
            fx = (C.c_float * len(flat))(*flat)
        rc = self._lib.fvvdp_b200_score_block(self.handle, tp, rp, st, int(n_frames), fx, C.c_void_p(q_ptr), int(q_stride),
                                              int(q_col0), C.c_void_p(flags_ptr) if flags_ptr else None, C.c_void_p(stream))
        self._check(rc, "fvvdp_b200_score_block")

    def score_block_yuv(self, desc, test_ptrs, ref_ptrs, n_frames, fixation_xy, q_ptr, q_stride, q_col0, stream):
        """score_block for device copies of raw planar Y'CbCr frames (file layout) of the window slots."""
        n = len(test_ptrs)
        assert n == len(ref_ptrs)
        tp = (C.c_void_p * n)(*test_ptrs)
        rp = (C.c_void_p * n)(*ref_ptrs)
        fx = None
        if fixation_xy is not None:
            flat = [float(v) for xy in fixation_xy for v in xy]
            fx = (C.c_float * len(flat))(*flat)
        self._lib.fvvdp_b200_score_block_yuv.restype = C.c_int
        rc = self._lib.fvvdp_b200_score_block_yuv(self.handle, C.byref(desc), tp, rp, C.c_int(int(n_frames)), fx, C.c_void_p(q_ptr),
                                                  C.c_int64(int(q_stride)), C.c_int64(int(q_col0)), C.c_void_p(stream))
        self._check(rc, "fvvdp_b200_score_block_yuv")

    def heatmap(self, frame, beta_jod, jod_a_abs, out_ptr, stream):
        self._check(self._lib.fvvdp_b200_heatmap(self.handle, int(frame), float(beta_jod), float(jod_a_abs), C.c_void_p(out_ptr),
                                                 C.c_void_p(stream)), "fvvdp_b200_heatmap")

    def heatmap_visualize(self, frame, beta_jod, jod_a_abs, colormap, out_ptr, stream):
        self._check(self._lib.fvvdp_b200_heatmap_visualize(self.handle, int(frame), float(beta_jod), float(jod_a_abs), COLORMAPS[colormap],
                                                           C.c_void_p(out_ptr), C.c_void_p(stream)), "fvvdp_b200_heatmap_visualize")

    def set_foveation_maps(self, level, view_ptr, log2_rho_ptr):
        self._check(self._lib.fvvdp_b200_set_foveation_maps(self.handle, int(level), C.c_void_p(view_ptr), C.c_void_p(log2_rho_ptr)),
                    "fvvdp_b200_set_foveation_maps")

    def read_tap(self, tap, level, frame, dst_ptr, capacity, stream):
        return self._check(self._lib.fvvdp_b200_read_tap(self.handle, int(tap), int(level), int(frame), C.c_void_p(dst_ptr),
                                                         int(capacity), C.c_void_p(stream)), "fvvdp_b200_read_tap")

    def level_size(self, level):
        h, w = C.c_int32(), C.c_int32()
        self._check(self._lib.fvvdp_b200_level_size(self.handle, int(level), C.byref(h), C.byref(w)), "fvvdp_b200_level_size")
        return h.value, w.value

    def profile(self, enable):
        self._check(self._lib.fvvdp_b200_profile(self.handle, 1 if enable else 0), "fvvdp_b200_profile")

    def profile_read(self):
        """{class name: (milliseconds, launches)} since the last read; classes 'front', 'level0'.., 'final'."""
        ms = (C.c_float * PROFILE_CLASSES)()
        cnt = (C.c_int32 * PROFILE_CLASSES)()
        self._check(self._lib.fvvdp_b200_profile_read(self.handle, ms, cnt), "fvvdp_b200_profile_read")
        names = ["front"] + [f"level{l}" for l in range(MAX_LEVELS)] + ["final"]
        return {names[i]: (float(ms[i]), int(cnt[i])) for i in range(PROFILE_CLASSES) if cnt[i] > 0}

    def launch_count(self):
        return int(self._lib.fvvdp_b200_launch_count(self.handle))

    def traffic_model(self):
        out = (C.c_double * 2)()
        self._check(self._lib.fvvdp_b200_traffic_model(self.handle, out), "fvvdp_b200_traffic_model")
        return float(out[0]), float(out[1])


def yuv_to_luminance(desc: YuvDesc, y_ptr, u_ptr, v_ptr, lum_ptr, rgb_ptr, device_index, stream):
    """One planar Y'CbCr frame (device planes) -> luminance (H,W) and / or display-encoded RGB (H,W,3)."""
    lib = load_library()
    rc = lib.fvvdp_b200_yuv_to_luminance(C.byref(desc), C.c_void_p(y_ptr), C.c_void_p(u_ptr), C.c_void_p(v_ptr),
                                         C.c_void_p(lum_ptr) if lum_ptr else None, C.c_void_p(rgb_ptr) if rgb_ptr else None,
                                         int(device_index), C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"fvvdp_b200_yuv_to_luminance failed ({rc}): {last_error(None)}")


def pu_sq_err(test_ptr, ref_ptr, n, params: PuParams, acc_ptr, device_index, stream):
    """acc (device double) += sum((PU(test) - PU(ref))^2) over n luminance samples."""
    lib = load_library()
    rc = lib.fvvdp_b200_pu_sq_err(C.c_void_p(test_ptr), C.c_void_p(ref_ptr), int(n), C.byref(params), C.c_void_p(acc_ptr), int(device_index),
                                  C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"fvvdp_b200_pu_sq_err failed ({rc}): {last_error(None)}")


class FrameFormat(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("in_dtype", C.c_int32), ("in_channels", C.c_int32), ("eotf", C.c_int32),
                ("Y_peak", C.c_float), ("Y_black", C.c_float), ("gamma", C.c_float), ("L_min", C.c_float), ("L_max", C.c_float),
                ("rgb2y", C.c_float * 3)]


def pu_sq_err_frames(fmt: FrameFormat, test_ptrs, ref_ptrs, strides, params: PuParams, out_ptr, device_index, stream):
    """out[i] (device double) += sum((PU(test_i) - PU(ref_i))^2) for a block of frames in their own dtype / layout."""
    lib = load_library()
    n = len(test_ptrs)
    tp = (C.c_void_p * n)(*test_ptrs)
    rp = (C.c_void_p * n)(*ref_ptrs)
    st = (C.c_int64 * 3)(*[int(v) for v in strides])
    lib.fvvdp_b200_pu_sq_err_frames.restype = C.c_int
    rc = lib.fvvdp_b200_pu_sq_err_frames(C.byref(fmt), tp, rp, st, C.c_int(n), C.byref(params), C.c_void_p(out_ptr), C.c_int(device_index), C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"fvvdp_b200_pu_sq_err_frames failed ({rc}): {last_error(None)}")


def pool_jod(q_ptr, n_bands, n_frames, q_stride, params: PoolParams, device_index, out_ptr, stream):
    """do_pooling_and_jods on the device: out[0] = JOD, out[1] = pooled Q."""
    lib = load_library()
    rc = lib.fvvdp_b200_pool_jod(C.c_void_p(q_ptr), int(n_bands), int(n_frames), int(q_stride), C.byref(params), int(device_index),
                                 C.c_void_p(out_ptr), C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"fvvdp_b200_pool_jod failed ({rc}): {last_error(None)}")
