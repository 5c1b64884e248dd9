"""The FovVideoVDP metric object, B200-native.

Public surface = the reference's `pyfvvdp.fvvdp` class (pyfvvdp/fvvdp.py:58-665): constructor arguments,
predict(), predict_video_source(), set_display_model(), update_device(), load_config(), get_info_string(),
short_name(), quality_unit(), write_features_to_json(), and the `stats` dictionary.  What differs is what
happens inside predict_video_source(): instead of ~7k tensor ops per frame (fvvdp.py:246-311,359-478) the
frames are scored in blocks by the hand-written sm_100a kernels of libfvvdp_b200.so (C ABI in
include/fvvdp_b200.h), called through ctypes with raw device pointers.

There is no CPU fallback: constructing the metric without a CUDA device, or with device='cpu', raises.
The metric is inference-only (`use_checkpoints` is accepted for signature compatibility and ignored).
"""
import ctypes
import json
import logging
import math
import threading

import numpy as np
import torch

from . import _native, config
from .display_model import (fvvdp_display_geometry, fvvdp_display_photometry, geometry_is_stock, photometry_kernel_spec)
from .video_source import fvvdp_video_source_array, is_array_source

_DTYPES = {torch.float32: _native.DTYPE_F32, torch.uint8: _native.DTYPE_U8, torch.int16: _native.DTYPE_U16}
_WORKSPACE_BUDGET_BYTES = 32e9  # device memory a scoring context may take for its per-block pyramids (4K: the full 96-frame block of the general kernels)
_HOST_BLOCK_FRAMES = 8          # clips in host memory: frames per block, so that uploads and kernels overlap block by block


_LAYOUT_CACHE, _FILTER_CACHE = {}, {}


def cached_pyramid_layout(width, height, ppd):
    key = (width, height, float(ppd))
    if key not in _LAYOUT_CACHE:
        _LAYOUT_CACHE[key] = pyramid_layout(width, height, ppd)
    return _LAYOUT_CACHE[key]


def cached_temporal_filters(frames_per_second, filter_len, sigma, beta):
    key = (float(frames_per_second), filter_len, float(sigma), float(beta))
    if key not in _FILTER_CACHE:
        F = temporal_filters(frames_per_second, filter_len, sigma, beta)
        _FILTER_CACHE[key] = (F, F.tobytes(), torch.from_numpy(F))
    return _FILTER_CACHE[key]


def pyramid_layout(width, height, ppd):
    """Number of Gaussian levels and the band centre frequencies [cpd] for a W x H frame seen at `ppd`
    pixels per degree (fvvdp_lpyr_dec.__init__, fvvdp_lpyr_dec.py:15-49).  Returns (n_levels, freqs) with
    len(freqs) == n_levels; the last entry belongs to the base band, which is never scored."""
    most = int(math.floor(math.log2(min(height, width)))) - 1
    # band k>=1 peaks at 0.3228 * 2^-(k-1) of the Nyquist frequency ppd/2; band 0 sits at Nyquist
    peak = [1.0] + [0.3228 * 2.0 ** (-k) for k in range(14)]
    n_bands = most
    for k, rel in enumerate(peak):
        if rel * ppd / 2.0 <= 0.5:  # bands at or below 0.5 cpd are folded into the base band
            n_bands = k
            break
    n_bands = max(0, min(n_bands + 1, most))
    freqs = np.array(peak[:n_bands + 1], dtype=np.float64) * ppd / 2.0
    return n_bands + 1, freqs


def temporal_filters(frames_per_second, filter_len, sigma, beta):
    """Sustained / transient temporal impulse responses sampled at the frame times, float32 (2, filter_len)
    (get_temporal_filters, fvvdp.py:609-630).  Tap 0 weighs the newest frame."""
    f32 = np.float32
    t = np.linspace(0.0, filter_len / frames_per_second, filter_len).astype(f32)
    lg = np.log(t + f32(1e-4)) - np.log(f32(beta))
    sust = np.exp(-(lg * lg) / f32(2.0 * sigma * sigma)).astype(f32)
    sust = sust / sust.sum(dtype=f32)
    trans = np.zeros(filter_len, f32)
    if filter_len > 1:
        trans[:-1] = f32(0.062170507756932) * (np.diff(sust) / (t[1] - t[0]))
    return np.stack([sust, trans]).astype(f32)


def initial_window(n_frames, filter_len, temp_padding):
    """Frame indices that fill the temporal window when frame 0 is scored, oldest first (fvvdp.py:258-285)."""
    fl = filter_len
    if temp_padding == "replicate":
        return [0] * fl
    if temp_padding == "circular":
        return [(n_frames - 1 - fl + k) % n_frames for k in range(fl)]
    if temp_padding == "pingpong":
        there_and_back = list(range(n_frames)) + list(range(n_frames - 2, 0, -1))
        seq = []
        while len(seq) < fl - 1:
            seq = seq + there_and_back
        return (seq[-(fl - 1):] if fl > 1 else []) + [0]
    raise RuntimeError('Unknown padding method "{}"'.format(temp_padding))


# a temporal-halo frame (staged, converted, reduced, filtered into the rings, not scored) relative to a scored one.  Measured with
# tools/halo_cost.py on the warp-specialised kernel (its producer warps do all of their work for a halo frame): 0.83-0.85; the
# cuts at 0.85-1.0 are the same for 2 and 8 ranks of 64 frames each and 1 % faster than those at 0.6
HALO_SLOT_COST = 0.9


def frame_block(n_frames, rank, world_size, halo=0, first_halo=None, halo_cost=HALO_SLOT_COST):
    """Contiguous block of frames [begin, end) scored by `rank` when a clip is sharded over `world_size` processes
    (SURVEY.md section 8e).  With `halo` > 0 the cut balances the WORK instead of the frame count: every rank but the first walks
    `halo` extra frames (the temporal window before its block) at `halo_cost` of a scored frame each, the first rank `first_halo`
    of them (1 with replicate padding, where the repeated first frame is reduced once), so the first rank takes a few frames
    more and all ranks reach the all-reduce together."""
    if halo <= 0 or world_size == 1:
        return (rank * n_frames) // world_size, ((rank + 1) * n_frames) // world_size
    first_halo = halo if first_halo is None else first_halo
    extra = [halo_cost * (first_halo if r == 0 else halo) for r in range(world_size)]
    share = (n_frames + sum(extra)) / world_size          # work units per rank
    bounds, acc = [0], 0.0
    for r in range(world_size - 1):
        acc += max(share - extra[r], 1.0)
        bounds.append(min(max(int(round(acc)), bounds[-1] + 1), n_frames - (world_size - 1 - r)))
    bounds.append(n_frames)
    return bounds[rank], bounds[rank + 1]


class _PinnedPool:
    """Small pool of pinned host buffers for the per-call result read-back (cudaHostAlloc is too slow to do per call)."""

    def __init__(self):
        self.free = {}
        self.lock = threading.Lock()

    def take(self, n):
        with self.lock:
            lst = self.free.get(n)
            if lst:
                return lst.pop()
        return torch.empty(n, dtype=torch.float32, pin_memory=True)

    def give(self, buf):
        with self.lock:
            lst = self.free.setdefault(buf.numel(), [])
            if len(lst) < 8:
                lst.append(buf)


_PINNED = _PinnedPool()


class _PinnedBytesPool:
    """Process-wide pool of pinned staging buffers for raw video frames, keyed by (elements, dtype)."""

    def __init__(self, cap_bytes=4 << 30):
        self.free, self.lock, self.cap, self.held = {}, threading.Lock(), cap_bytes, 0

    def take(self, n, dtype):
        with self.lock:
            lst = self.free.get((n, dtype))
            if lst:
                buf = lst.pop()
                self.held -= buf.numel() * buf.element_size()
                return buf
        return torch.empty(n, dtype=dtype, pin_memory=True)

    def give(self, buf):
        nbytes = buf.numel() * buf.element_size()
        with self.lock:
            if self.held + nbytes <= self.cap:
                self.free.setdefault((buf.numel(), buf.dtype), []).append(buf)
                self.held += nbytes


_PINNED_BYTES = _PinnedBytesPool()


class _LazyStats(dict):
    """The `stats` dictionary of predict(): "Q_per_ch" arrives from the device when the dictionary is first looked at."""

    def __init__(self, host, done, shape):
        super().__init__()
        self._pending = (host, done, shape)

    def _force(self):
        pend, self._pending = self._pending, None
        if pend is None:
            return
        host, done, shape = pend
        done.synchronize()
        arr = host.numpy().copy()
        _PINNED.give(host)
        dict.__setitem__(self, "Q_per_ch", arr[:-1].reshape(shape))
        if arr[-1:].view(np.uint32)[0] != 0:
            logging.warning("Pixel outside the valid range 0-1")

    def __getitem__(self, k):
        self._force()
        return dict.__getitem__(self, k)

    def __contains__(self, k):
        self._force()
        return dict.__contains__(self, k)

    def __iter__(self):
        self._force()
        return dict.__iter__(self)

    def __len__(self):
        self._force()
        return dict.__len__(self)

    def __repr__(self):
        self._force()
        return dict.__repr__(self)

    def get(self, k, default=None):
        self._force()
        return dict.get(self, k, default)

    def keys(self):
        self._force()
        return dict.keys(self)

    def items(self):
        self._force()
        return dict.items(self)

    def values(self):
        self._force()
        return dict.values(self)

    def copy(self):
        self._force()
        return dict(self)

    def __del__(self):
        try:
            pend = self._pending
            if pend is not None:
                pend[1].synchronize()   # the warning is still owed to the caller who never looked at the statistics
                if pend[0].numpy()[-1:].view(np.uint32)[0] != 0:
                    logging.warning("Pixel outside the valid range 0-1")
                _PINNED.give(pend[0])
        except Exception:
            pass


class _FrameSet:
    """Resolves frame index -> (test pointer, reference pointer) on the device for one clip.

    Three kinds of source:
      * array source already on the metric's device: pointers into the user's tensors, nothing is copied;
      * array source in host memory: frames are uploaded (raw dtype, original memory layout) into a small pool
        of device buffers as blocks need them;
      * any other fvvdp_video_source: get_test_frame()/get_reference_frame() supply float32 luminance frames.
    """

    def __init__(self, vid_source, device, raw):
        self.vs = vid_source
        self.device = device
        self.raw = raw
        self.held = {}      # frame index -> (test tensor, ref tensor) on the device
        self.free = []      # recycled upload buffers: (buffer, event after which the kernels no longer read it), oldest first
        self.h2d_bytes = 0
        self.up_stream = None   # host-resident array sources: uploads run on their own stream, one block ahead of the kernels
        if raw:
            tv, rv = vid_source.test_video, vid_source.reference_video
            if tv.shape[0] != 1:
                raise RuntimeError("Batches of more than one clip are not supported (the reference is limited to B=1 as well)")
            if tv.dtype != rv.dtype or tv.dtype not in _DTYPES:
                raise RuntimeError("Only uint8, uint16 and float32 is currently supported")
            self.dtype = tv.dtype
            self.resident = tv.device == device and rv.device == device
            self.same_layout = tv.stride() == rv.stride()
            v = tv[0, :, 0]
            if self.resident and self.same_layout:
                self.strides = tuple(v.stride())
                # frame k of a resident clip sits at base + k * frame stride: pointers by arithmetic, no tensor views
                self._base = (tv.data_ptr(), rv.data_ptr())
                self._frame_bytes = tv.stride(2) * tv.element_size()
                self._n_local = tv.shape[2]
            else:
                self._order = sorted(range(3), key=lambda d: -v.stride(d))
                vp = v.permute(self._order)
                if not (vp.is_contiguous() and self.same_layout):
                    self._order = [0, 1, 2]
                    vp = v
                self._buf_shape = tuple(vp.shape)
                inv = [self._order.index(d) for d in range(3)]
                self._inv = inv
                self.strides = tuple(torch.empty(self._buf_shape, dtype=self.dtype, device="meta").permute(inv).stride())
                self.resident = False
                self.up_stream = torch.cuda.Stream(device=device)
        else:
            self.dtype = torch.float32
            self.strides = None

    def _upload(self, src):
        if self.free:
            buf, reusable = self.free.pop(0)
            if reusable is not None:
                self.up_stream.wait_event(reusable)  # the block that read this buffer last has been scored
        else:
            buf = torch.empty(self._buf_shape, dtype=self.dtype, device=self.device)
            # the allocator may hand out memory that work queued on the current stream still reads
            self.up_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.up_stream):
            buf.copy_(src.permute(self._order), non_blocking=True)
        self.h2d_bytes += buf.numel() * buf.element_size()
        return buf

    def uploaded(self):
        """Event on the upload stream after everything fetched so far (None when nothing is uploaded asynchronously)."""
        if self.up_stream is None:
            return None
        ev = torch.cuda.Event()
        ev.record(self.up_stream)
        return ev

    def fetch_all(self, indices):
        """fetch() for every frame index of a block."""
        if self.raw and self.resident:  # nothing to move: range check only
            first = getattr(self.vs, "first_frame", 0)
            lo, hi = min(indices) - first, max(indices) - first
            if lo < 0 or hi >= self._n_local:
                raise RuntimeError(f"frames {min(indices)}..{max(indices)} are not all held by this process")
            return
        for idx in indices:
            self.fetch(idx)

    def block_pointers(self, indices):
        """([test pointers], [reference pointers]) of the frames of a block."""
        if self.raw and self.resident:
            first, fb, (bt, br) = getattr(self.vs, "first_frame", 0), self._frame_bytes, self._base
            offs = [(i - first) * fb for i in indices]
            return [bt + o for o in offs], [br + o for o in offs]
        ptrs = [self.pointers(i) for i in indices]
        return [q[0] for q in ptrs], [q[1] for q in ptrs]

    def fetch(self, idx):
        if idx in self.held:
            return
        if self.raw:
            tv, rv = self.vs.test_video, self.vs.reference_video
            k = self.vs.local_index(idx) if hasattr(self.vs, "local_index") else idx
            if self.resident:
                if k < 0 or k >= self._n_local:
                    raise RuntimeError(f"frame {idx} is not held by this process")
                self.held[idx] = k
            else:
                self.held[idx] = (self._upload(tv[0, :, k]), self._upload(rv[0, :, k]))
        else:
            t = self.vs.get_test_frame(idx, device=self.device)
            r = self.vs.get_reference_frame(idx, device=self.device)
            t = t.to(device=self.device, dtype=torch.float32).reshape(t.shape[-2], t.shape[-1]).contiguous()
            r = r.to(device=self.device, dtype=torch.float32).reshape(r.shape[-2], r.shape[-1]).contiguous()
            self.held[idx] = (t, r)
            if self.strides is None:
                self.strides = (0, t.stride(0), t.stride(1))

    def pointers(self, idx):
        if self.raw and self.resident:
            off = self.held[idx] * self._frame_bytes
            return self._base[0] + off, self._base[1] + off
        t, r = self.held[idx]
        return t.data_ptr(), r.data_ptr()

    def retain_only(self, keep, scored=None):
        """Drop every held frame that is not in `keep`; `scored` = event after the kernels that read them."""
        if self.raw and self.resident:
            self.held.clear()
            return
        for idx in [k for k in self.held if k not in keep]:
            t, r = self.held.pop(idx)
            if self.raw and not self.resident:
                self.free.extend(((t, scored), (r, scored)))


class _YuvFrames:
    """Device copies of the raw frames of a fvvdp_video_source_yuv_file (file layout: Y plane, Cb plane, Cr plane), uploaded
    block by block: worker threads copy the mem-mapped frames into pinned staging buffers, a copy stream moves them to the
    device one block ahead of the kernels.  Same interface as _FrameSet as far as predict_video_source() uses it."""

    def __init__(self, vid_source, device):
        from concurrent.futures import ThreadPoolExecutor
        self.readers = (vid_source.test_vidr, vid_source.reference_vidr)
        self.device = device
        self.raw, self.resident, self.strides = False, False, None
        self.held, self.free, self.pinned = {}, [], []
        self.h2d_bytes = 0
        self.up_stream = torch.cuda.Stream(device=device)
        self.pool = ThreadPoolExecutor(max_workers=8)
        for rd in self.readers:
            if rd.mm is None:
                rd.mm = np.memmap(rd.file_name, rd.dtype, mode="r")
        self.tdtype = torch.int16 if self.readers[0].dtype == np.uint16 else torch.uint8

    def __del__(self):  # the staging buffers go back to the process-wide pool (pinning memory is far slower than copying into it)
        try:
            for buf, ev in self.pinned:
                if ev is not None:
                    ev.synchronize()
                _PINNED_BYTES.give(buf)
        except Exception:
            pass

    def _stage(self, which, idx):
        """mem-mapped frame -> pinned staging buffer (runs in a worker thread; numpy releases the GIL for the copy)."""
        reader = self.readers[which]
        n = reader.frame_pixels
        if self.pinned:
            buf, ev = self.pinned.pop()
            if ev is not None:
                ev.synchronize()  # its previous upload has left the buffer
        else:
            buf = _PINNED_BYTES.take(n, self.tdtype)
        o = int(idx) * n
        np.copyto(buf.numpy().view(reader.dtype), reader.mm[o:o + n])
        return buf

    def fetch_all(self, indices):
        todo = [i for i in dict.fromkeys(indices) if i not in self.held]
        if not todo:
            return
        for i in todo:
            if i < 0 or i >= self.readers[0].frame_count or i >= self.readers[1].frame_count:
                raise RuntimeError("The frame index is outside the range of available frames")
        staged = list(self.pool.map(lambda job: self._stage(*job), [(w, i) for i in todo for w in range(2)]))
        with torch.cuda.stream(self.up_stream):
            for k, i in enumerate(todo):
                pair = []
                for st in range(2):
                    host = staged[2 * k + st]
                    if self.free:
                        dev, reusable = self.free.pop(0)
                        if reusable is not None:
                            self.up_stream.wait_event(reusable)
                    else:
                        dev = torch.empty(host.numel(), dtype=self.tdtype, device=self.device)
                        self.up_stream.wait_stream(torch.cuda.current_stream(self.device))
                    dev.copy_(host, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.up_stream)
                    self.pinned.append((host, ev))
                    self.h2d_bytes += host.numel() * host.element_size()
                    pair.append(dev)
                self.held[i] = tuple(pair)

    def uploaded(self):
        ev = torch.cuda.Event()
        ev.record(self.up_stream)
        return ev

    def block_pointers(self, indices):
        return [self.held[i][0].data_ptr() for i in indices], [self.held[i][1].data_ptr() for i in indices]

    def retain_only(self, keep, scored=None):
        for idx in [k for k in self.held if k not in keep]:
            t, r = self.held.pop(idx)
            self.free.extend(((t, scored), (r, scored)))


class fvvdp:
    def __init__(self, display_name="standard_4k", display_photometry=None, display_geometry=None, color_space="sRGB", foveated=False,
                 heatmap=None, quiet=False, device=None, temp_padding="replicate", use_checkpoints=False, block_frames=None,
                 shard_frames=False):
        assert heatmap in [None, "none", "raw", "threshold", "supra-threshold"], "Unsupported heatmap type"
        assert temp_padding in ["replicate", "circular", "pingpong"], "Unsupported temporal padding method"
        self.quiet = quiet
        self.foveated = foveated
        self.heatmap = heatmap
        self.color_space = color_space
        self.temp_padding = temp_padding
        self.use_checkpoints = use_checkpoints
        self.do_heatmap = heatmap is not None and heatmap != "none"
        self.block_frames = block_frames
        self.shard_frames = shard_frames
        self.debug_taps = False           # tests: keep intermediate tensors of the last block readable
        self._ctx = None
        self._ctx_key = None
        self.last_run = {}                # launch / traffic counters of the last predict call (bench.py)

        if device is None:
            if not (torch.cuda.is_available() and torch.cuda.device_count() > 0):
                raise RuntimeError("fovvideovdp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda:0")
        self.update_device(device, _reload=False)
        _native.load_library()  # fail now, loudly, if the CUDA library is missing
        self.set_display_model(display_name, display_photometry=display_photometry, display_geometry=display_geometry)
        self.load_config()
        self.lut = config.csf_lut()

    # ------------------------------------------------------------------ configuration
    def update_device(self, device, _reload=True):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"fovvideovdp_b200 runs on CUDA devices only (got '{device}'); there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self._drop_ctx()

    def load_config(self):
        p = config.parameters()
        self.parameters_file = config.config_files.find("fvvdp_parameters.json") or "<packaged>"
        for k in ("mask_p", "mask_c", "pu_dilate", "w_transient", "beta", "beta_t", "beta_tch", "beta_sch", "sustained_sigma",
                  "sustained_beta", "csf_sigma", "sensitivity_correction", "masking_model", "local_adapt", "contrast", "jod_a",
                  "log_jod_exp", "mask_q_sust", "mask_q_trans", "k_cm", "filter_len", "version"):
            setattr(self, k, p[k])
        # the kernels implement the shipped calibration's model structure
        if self.local_adapt != "gpyr" or self.contrast != "weber" or self.masking_model != "min_mutual_masking_perc_norm2" or self.pu_dilate != 0:
            raise RuntimeError("Unsupported metric configuration: the B200 core implements local_adapt='gpyr', contrast='weber', "
                               "masking_model='min_mutual_masking_perc_norm2', pu_dilate=0")
        # the packaged CSF look-up tables were computed for one (csf_sigma, k_cm) pair; the reference keys its cache files on
        # them (fvvdp.py:502-518) and fails when no file matches -- never score with the wrong sensitivity table
        if abs(self.csf_sigma - config.CSF_LUT_KEY["csf_sigma"]) > 1e-6 or abs(self.k_cm - config.CSF_LUT_KEY["k_cm"]) > 1e-6:
            raise RuntimeError("CSF look-up table for csf_sigma={}, k_cm={} not found (the packaged table is for csf_sigma={}, k_cm={})".format(
                self.csf_sigma, self.k_cm, config.CSF_LUT_KEY["csf_sigma"], config.CSF_LUT_KEY["k_cm"]))
        self._drop_ctx()

    def set_display_model(self, display_name="standard_4k", display_photometry=None, display_geometry=None):
        if display_photometry is None:
            self.display_photometry = fvvdp_display_photometry.load(display_name)
            self.display_name = display_name
        else:
            self.display_photometry = display_photometry
            self.display_name = "unspecified"
        self.display_geometry = fvvdp_display_geometry.load(display_name) if display_geometry is None else display_geometry
        self.pix_per_deg = self.display_geometry.get_ppd()
        self._drop_ctx()

    def _drop_ctx(self):
        if getattr(self, "_ctx", None) is not None:
            self._ctx.close()
        self._ctx = None
        self._ctx_key = None

    # ------------------------------------------------------------------ prediction
    def predict(self, test_cont, reference_cont, dim_order="BCFHW", frames_per_second=0, fixation_point=None):
        vs = fvvdp_video_source_array(test_cont, reference_cont, frames_per_second, dim_order=dim_order,
                                      display_photometry=self.display_photometry, color_space_name=self.color_space)
        return self.predict_video_source(vs, fixation_point=fixation_point)

    def _context(self, key, build_cfg):
        if self._ctx is None or self._ctx_key != key:
            self._drop_ctx()
            cfg, keep = build_cfg()
            self._ctx = _native.Context(cfg, self.device.index, keepalive=keep)
            self._ctx_key = key
        return self._ctx

    def _make_config(self, W, H, n_levels, freqs, temp_ch, fl, F, spec, dtype, C, T, rgb2y):
        cfg = _native.Config()
        cfg.abi_version = _native.ABI_VERSION
        cfg.width, cfg.height, cfg.n_levels = W, H, n_levels
        for i, f in enumerate(freqs):
            cfg.band_freq[i] = float(f)
        cfg.temp_ch, cfg.filter_len = temp_ch, fl
        for cc in range(F.shape[0]):
            for k in range(fl):
                cfg.filt[cc][k] = float(F[cc, k])
        cfg.eotf = _native.EOTF_CODES[spec["kind"]]
        cfg.Y_peak = spec.get("Y_peak", 0.0)
        cfg.Y_black = spec.get("Y_black", 0.0)
        cfg.gamma = spec.get("gamma", 2.2)
        cfg.L_min = spec.get("L_min", 0.0)
        cfg.L_max = spec.get("L_max", 0.0)
        w = rgb2y if C == 3 else [1.0, 0.0, 0.0]
        for i in range(3):
            cfg.rgb2y[i] = w[i]
        cfg.in_dtype, cfg.in_channels = _DTYPES[dtype], C
        lut = self.lut
        keep = [np.ascontiguousarray(lut[k], np.float32) for k in ("rho_log", "Y_log", "ecc_sqrt", "S_log")]
        cfg.csf_rho_log, cfg.csf_Y_log, cfg.csf_ecc_sqrt, cfg.csf_S_log = [a.ctypes.data_as(ctypes.c_void_p) for a in keep]
        for name, dst in (("rho", cfg.csf_rho_range), ("Y", cfg.csf_Y_range), ("ecc", cfg.csf_ecc_range)):
            dst[0], dst[1] = float(lut[name][0]), float(lut[name][-1])
        cfg.mask_p = self.mask_p
        cfg.mask_q[0], cfg.mask_q[1] = self.mask_q_sust, self.mask_q_trans
        cfg.mask_c_mul = 10.0 ** self.mask_c
        cfg.sens_mul = 10.0 ** (self.sensitivity_correction / 20.0)
        cfg.beta = self.beta
        cfg.w_transient = self.w_transient
        geo = self.display_geometry
        cfg.foveated = 0
        if self.foveated:
            # a fvvdp_display_geometry subclass supplies its own per-band maps (fvvdp_b200_set_foveation_maps)
            cfg.foveated = 1 if geometry_is_stock(geo) else 2
            if cfg.foveated == 1:
                cfg.display_size_m[0], cfg.display_size_m[1] = float(geo.display_size_m[0]), float(geo.display_size_m[1])
                cfg.distance_m = float(geo.distance_m)
        cfg.ppd_centre = float(self.pix_per_deg)
        cfg.want_dmap = 0 if not self.do_heatmap else (1 if self.heatmap == "raw" else 2)
        cfg.want_taps = 1 if self.debug_taps else 0
        cfg.max_block_frames = T
        return cfg, keep

    def predict_video_source(self, vid_source, fixation_point=None):
        height, width, N_frames = vid_source.get_video_size()
        height, width, N_frames = int(height), int(width), int(N_frames)
        dev = self.device
        is_image = N_frames == 1
        fps = vid_source.get_frames_per_second()

        if fixation_point is None:
            fixation_point = np.array([width // 2, height // 2], dtype=np.float32)
        elif torch.is_tensor(fixation_point):
            fixation_point = fixation_point.detach().cpu().numpy()
        fixation_point = np.asarray(fixation_point, dtype=np.float32)

        n_levels, freqs = cached_pyramid_layout(width, height, self.pix_per_deg)
        n_bands = n_levels - 1
        if n_bands < 1:
            raise RuntimeError(f"Frames of {width}x{height} are too small to build a contrast pyramid")
        if is_image:
            temp_ch, fl = 1, 1
            F = np.ones((1, 1), np.float32)
            F_bytes = F.tobytes()
        else:
            temp_ch = 2
            fl = int(math.ceil(250.0 / (1000.0 / fps)))
            self.filter_len = fl
            if fl > _native.MAX_FILTER_LEN:
                raise RuntimeError(f"frame rate {fps} needs {fl} filter taps; at most {_native.MAX_FILTER_LEN} are supported")
            F, F_bytes, self.F = cached_temporal_filters(fps, fl, self.sustained_sigma, self.sustained_beta)

        # how the frames reach the kernels
        spec = None
        if is_array_source(vid_source):
            spec = photometry_kernel_spec(vid_source.dm_photometry)
        raw = spec is not None
        # raw .yuv clips with a stock display model: the frames go to the device as stored and one kernel per block converts them
        # into the planes level 0 stages (fvvdp_b200_score_block_yuv), full-screen resize included; custom photometry: get_*_frame()
        yuv_desc = None
        if (type(vid_source).__name__ == "fvvdp_video_source_yuv_file" and type(vid_source).__module__.startswith("fovvideovdp_b200")
                and getattr(vid_source, "_spec", None) is not None and not is_image
                and fl <= 16 and not self.debug_taps and not self.do_heatmap):
            tr, rr = vid_source.test_vidr, vid_source.reference_vidr
            if (tr.width, tr.height, tr.bit_depth, tr.chroma_ss, tr.color_space) == (rr.width, rr.height, rr.bit_depth, rr.chroma_ss, rr.color_space):
                yuv_desc = tr._desc(vid_source._spec, vid_source.color_to_luminance, vid_source.resize_of(tr))
        frames = _YuvFrames(vid_source, dev) if yuv_desc is not None else _FrameSet(vid_source, dev, raw)
        if raw:
            C = 3 if vid_source.is_color else 1
            dtype = frames.dtype
            # the RGB -> luminance weights are the SOURCE's (video_source.py:87,206), not the metric's colour space
            rgb2y = tuple(float(v) for v in vid_source.color_to_luminance)
        else:
            spec, C, dtype = dict(kind="none"), 1, torch.float32
            rgb2y = (1.0, 0.0, 0.0)

        # frames this process scores
        f_begin, f_end = 0, N_frames
        world = 1
        if self.shard_frames and torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            if world > 1 and N_frames >= world:
                f_begin, f_end = frame_block(N_frames, torch.distributed.get_rank(), world, halo=fl - 1,
                                             first_halo=1 if self.temp_padding == "replicate" else fl - 1)
            else:
                world = 1  # replicas only: every rank scores the whole (short) clip

        per_frame_bytes = 4.0 * (2 * temp_ch) * height * width * 4.0 / 3.0 * (3.0 if self.debug_taps else 1.0)
        T = self.block_frames or int(max(1, min(_native.MAX_BLOCK_FRAMES, _WORKSPACE_BUDGET_BYTES // per_frame_bytes)))
        if frames.up_stream is not None and not self.block_frames and not self.debug_taps:
            T = min(T, _HOST_BLOCK_FRAMES)
        T = max(1, min(T, _native.MAX_BLOCK_FRAMES, f_end - f_begin))

        geo = self.display_geometry
        key = (width, height, n_levels, float(self.pix_per_deg), temp_ch, fl, F_bytes, tuple(sorted(spec.items())), dtype, C, T,
               self.foveated, self.heatmap if self.do_heatmap else None, self.debug_taps, rgb2y,
               (tuple(geo.display_size_m), geo.distance_m, geometry_is_stock(geo)) if self.foveated else None)
        ctx = self._context(key, lambda: self._make_config(width, height, n_levels, freqs, temp_ch, fl, F, spec, dtype, C, T, rgb2y))

        stream = torch.cuda.current_stream(dev).cuda_stream
        # one buffer: the pooled energies and, behind them, the flag word -- one allocation, one device->host read
        result = torch.zeros(n_bands * 2 * N_frames + 1, dtype=torch.float32, device=dev)
        Q_per_ch = result[:-1].view(n_bands, 2, N_frames)
        flags = result[-1:].view(torch.int32)
        heatmap = None
        if self.do_heatmap:
            hm_ch = 1 if self.heatmap == "raw" else 3  # fvvdp.py:236
            heatmap = torch.zeros([1, hm_ch, N_frames, height, width], dtype=torch.float16, device="cpu")
            # a ring of device / pinned staging planes: the device->host copy of frame j overlaps the kernels of frames j+1..j+3,
            # the host copies a plane into the (pageable) result when its event has fired
            HM_RING = 4
            hm_dev = [torch.empty((hm_ch, height, width), dtype=torch.float16, device=dev) for _ in range(HM_RING)]
            hm_pin = [_PINNED_BYTES.take(hm_ch * height * width, torch.float16).view(hm_ch, height, width) for _ in range(HM_RING)]
            hm_ev, hm_frame, hm_count = [None] * HM_RING, [None] * HM_RING, 0

            def hm_collect(k):
                if hm_frame[k] is not None:
                    hm_ev[k].synchronize()
                    heatmap[0, :, hm_frame[k]].copy_(hm_pin[k])
                    hm_frame[k] = None
        custom_geo = self.foveated and not geometry_is_stock(geo)
        if custom_geo:
            fov_maps = self._custom_foveation_maps(ctx, n_bands, freqs)  # kept alive until the end of this call
        first = initial_window(N_frames, fl, self.temp_padding) if not is_image else [0]

        # frame shown at time t = -(fl-1) .. N-1 (t <= 0 falls into the temporal padding): timeline[t + fl - 1]
        timeline = list(first[:fl]) + list(range(1, N_frames))

        def block_slots(f0):
            n = min(T, f_end - f0)
            return n, timeline[f0:f0 + n + fl - 1]

        def prefetch(f0):  # frames of the block starting at f0 -> device (host sources: on the upload stream)
            frames.fetch_all(block_slots(f0)[1])
            return frames.uploaded()

        launches0 = ctx.launch_count()
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            ahead = frames.up_stream is not None
            ready = prefetch(f_begin) if ahead else None
            for f0 in range(f_begin, f_end, T):
                n, slots = block_slots(f0)
                nxt = f0 + n
                if not ahead:
                    prefetch(f0)
                # host-resident clips: the next block's uploads are queued before this block's kernels; they only take
                # buffers that were released before this block, so they run while this block is scored
                ready_next = prefetch(nxt) if (ahead and nxt < f_end) else None
                if ready is not None:
                    cur.wait_event(ready)
                test_ptrs, ref_ptrs = frames.block_pointers(slots)
                fix = None
                if self.foveated:
                    fix = [fixation_point[f0 + i] if fixation_point.ndim == 2 else fixation_point for i in range(n)]
                    if custom_geo:  # gaze direction [deg] through the plugin (fvvdp.py:429-431)
                        fix = [self._gaze_direction(xy, width, height) for xy in fix]
                if yuv_desc is not None:
                    ctx.score_block_yuv(yuv_desc, test_ptrs, ref_ptrs, n, fix, Q_per_ch.data_ptr(), N_frames, f0, stream)
                else:
                    ctx.score_block(test_ptrs, ref_ptrs, frames.strides, n, fix, Q_per_ch.data_ptr(), N_frames, f0,
                                    flags.data_ptr(), stream)
                if self.do_heatmap:
                    beta_jod = 10.0 ** self.log_jod_exp
                    for i in range(n):
                        k = hm_count % HM_RING
                        hm_count += 1
                        hm_collect(k)
                        if self.heatmap == "raw":
                            ctx.heatmap(i, beta_jod, abs(self.jod_a), hm_dev[k].data_ptr(), stream)
                        else:
                            ctx.heatmap_visualize(i, beta_jod, abs(self.jod_a), self.heatmap, hm_dev[k].data_ptr(), stream)
                        hm_pin[k].copy_(hm_dev[k], non_blocking=True)
                        hm_ev[k] = torch.cuda.Event()
                        hm_ev[k].record(cur)
                        hm_frame[k] = f0 + i
                scored = None
                if frames.up_stream is not None:
                    scored = torch.cuda.Event()
                    scored.record(cur)
                frames.retain_only(set(block_slots(nxt)[1]) if nxt < f_end else set(), scored)
                ready = ready_next
            if world > 1:
                torch.distributed.all_reduce(Q_per_ch)  # every column has exactly one non-zero contributor
            out = torch.empty(2, dtype=torch.float32, device=dev)
            pp = _native.PoolParams(self.beta_sch, self.beta_tch, self.beta_t, self.w_transient, self.jod_a, self.log_jod_exp)
            _native.pool_jod(Q_per_ch.data_ptr(), n_bands, N_frames, N_frames, pp, dev.index, out.data_ptr(), stream)

        alg, plan = ctx.traffic_model()
        self.last_run = dict(gpu_launches=ctx.launch_count() - launches0 + 1, block_frames=T, h2d_bytes=frames.h2d_bytes,
                             bytes_algorithmic_last_block=alg, bytes_plan_last_block=plan, frames_scored=f_end - f_begin)

        # The one device->host read of the call is an asynchronous copy into pinned memory: like the reference, predict returns a
        # device tensor for the JOD without waiting for the GPU; stats["Q_per_ch"] (and the out-of-range warning) materialise
        # when the dictionary is first looked at.  Back-to-back calls therefore keep the GPU busy across call boundaries.
        with torch.cuda.device(dev):
            host = _PINNED.take(result.numel())
            host.copy_(result, non_blocking=True)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(dev))
        stats = _LazyStats(host, done, (n_bands, 2, N_frames))
        stats["rho_band"] = freqs
        stats["frames_per_second"] = fps
        stats["width"] = width
        stats["height"] = height
        stats["N_frames"] = N_frames
        if self.do_heatmap:
            for k in range(HM_RING):
                hm_collect(k)
                _PINNED_BYTES.give(hm_pin[k].view(-1))
            stats["heatmap"] = heatmap
        return out[0], stats

    # ------------------------------------------------------------------ custom display geometry (foveated)
    def _custom_foveation_maps(self, ctx, n_bands, freqs):
        """Per-band view-direction and log2(rho) maps from the plugin's own pix2view_direction() and
        get_resolution_magnification(), as the reference calls them for every band (fvvdp.py:422-438)."""
        geo, dev, lut = self.display_geometry, self.device, self.lut
        keep = []
        for bb in range(n_bands):
            h, w = ctx.level_size(bb)
            xv = torch.linspace(0.5, w - 0.5, w, device=dev)
            yv = torch.linspace(0.5, h - 0.5, h, device=dev)
            xx, yy = torch.meshgrid(xv, yv, indexing="xy")
            view = geo.pix2view_direction(torch.tensor((w, h)), xx, yy).to(device=dev, dtype=torch.float32)
            res_mag = torch.as_tensor(geo.get_resolution_magnification(view), dtype=torch.float32, device=dev)
            rho = (float(freqs[bb]) * res_mag).expand(h, w)
            rq = torch.log2(torch.clamp(rho, float(lut["rho"][0]), float(lut["rho"][-1]))).contiguous()
            view = view.reshape(2, h, w).contiguous()
            ctx.set_foveation_maps(bb, view.data_ptr(), rq.data_ptr())
            keep.append((view, rq))
        return keep

    def _gaze_direction(self, xy, width, height):
        d = self.display_geometry.pix2view_direction(torch.tensor((width, height)), torch.as_tensor(float(xy[0]) + 0.5),
                                                     torch.as_tensor(float(xy[1]) + 0.5))
        return [float(d[0]), float(d[1])]

    # ------------------------------------------------------------------ debugging taps (tests)
    def read_tap(self, tap, level, frame_in_block):
        """Intermediate tensor of the last scored block (needs debug_taps=True before predict)."""
        ctx = self._ctx
        h, w = ctx.level_size(0 if tap == _native.TAP_R else level)
        nch = 2 * ctx.cfg.temp_ch
        planes = {_native.TAP_R: nch, _native.TAP_GAUSS: nch, _native.TAP_CONTRAST: nch, _native.TAP_LBKG: 1,
                  _native.TAP_S: ctx.cfg.temp_ch, _native.TAP_D: ctx.cfg.temp_ch, _native.TAP_DMAP_BAND: 1}[tap]
        dst = torch.empty((planes, h, w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ctx.read_tap(tap, level, frame_in_block, dst.data_ptr(), dst.numel(), torch.cuda.current_stream(self.device).cuda_stream)
        return dst

    # ------------------------------------------------------------------ reporting
    def short_name(self):
        return "FovVideoVDP"

    def quality_unit(self):
        return "JOD"

    def get_info_string(self):
        std = ", (" + self.display_name + ")" if self.display_name.startswith("standard_") else ""
        mode = "foveated" if self.foveated else "non-foveated"
        return '"FovVideoVDP v{}, {:.4g} [pix/deg], Lpeak={:.5g}, Lblack={:.4g} [cd/m^2], {}{}"'.format(
            self.version, self.pix_per_deg, self.display_photometry.get_peak_luminance(), self.display_photometry.get_black_level(), mode, std)

    def write_features_to_json(self, stats, dest_fname):
        Q = stats["Q_per_ch"]
        fmap = {}
        for k, v in stats.items():
            if k in ("Q_per_ch", "heatmap"):
                continue
            fmap[k] = v.tolist() if isinstance(v, np.ndarray) else v
        for cc in range(Q.shape[1]):
            for bb in range(Q.shape[0]):
                fmap[f"t{cc}_b{bb}"] = Q[bb, cc, :].tolist()
        with open(dest_fname, "w", encoding="utf-8") as f:
            json.dump(fmap, f, ensure_ascii=False, indent=4)
