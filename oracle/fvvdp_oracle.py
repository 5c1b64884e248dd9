"""CPU ORACLE for the FovVideoVDP per-frame hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a from-the-formulas numpy (float32) restatement of the reference's algorithm
(gfxdisp/FovVideoVDP, pyfvvdp 1.2.2 / parameters 1.2.3).  Each function cites the reference
file:line it follows.  It is written with explicit index arithmetic (mirror / replicate index
tables) instead of the reference's conv2d + additive fix-ups so that it is an independent
statement of the same maths.

Parity status: PINNED.  tests/test_oracle_golden.py checks it against tests/golden/*.npz, which were
produced by running the UNMODIFIED reference (torch CPU, fp32) in the build container with
tools/gen_golden.py, and against the README known answer (wavy_facade blur sigma=2 -> 8.693 JOD).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (fovvideovdp_b200/) never does; it fails loudly without its CUDA
library.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

_F = np.float32
_DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fovvideovdp_b200", "data")


# ----------------------------------------------------------------------------------------------
# data (calibration constants, display presets, CSF LUTs) -- read-only data files converted from the
# reference's data by tools/import_reference_data.py
# ----------------------------------------------------------------------------------------------
_cache = {}


def metric_data():
    if "md" not in _cache:
        with open(os.path.join(_DATA_DIR, "metric_data.json")) as f:
            _cache["md"] = json.load(f)
    return _cache["md"]


def csf_lut():
    """fvvdp.py:505-518 preload_cache: axes (32,) and S_log[omega_idx][Y][rho][ecc]."""
    if "lut" not in _cache:
        d = np.load(os.path.join(_DATA_DIR, "csf_lut.npz"))
        _cache["lut"] = {k: np.ascontiguousarray(d[k], dtype=_F) for k in d.files}
    return _cache["lut"]


# ----------------------------------------------------------------------------------------------
# display model (fvvdp_display_model.py)
# ----------------------------------------------------------------------------------------------
def photometry_from_preset(name):
    """fvvdp_display_photometry.load, fvvdp_display_model.py:49-98."""
    m = metric_data()["displays"][name]
    Y_peak = m["max_luminance"]
    if "min_luminance" in m:
        contrast = Y_peak / m["min_luminance"]
    else:
        contrast = m.get("contrast", 500)
    E_ambient = m.get("E_ambient", 0)
    k_refl = m.get("k_refl", 0.005)
    # get_black_level, fvvdp_display_model.py:172-176
    Y_black = E_ambient / math.pi * k_refl + Y_peak / contrast
    return dict(kind=m.get("EOTF", "sRGB"), Y_peak=float(Y_peak), Y_black=float(Y_black), gamma=float(m.get("gamma", 2.2)))


def geometry_from_preset(name):
    """fvvdp_display_geometry.load + __init__, fvvdp_display_model.py:385-436,542-568."""
    m = metric_data()["displays"][name]
    W, H = m["resolution"]
    if "viewing_distance_meters" in m:
        dist = m["viewing_distance_meters"]
    elif "viewing_distance_inches" in m:
        dist = m["viewing_distance_inches"] * 0.0254
    else:
        dist = None
    if "diagonal_size_meters" in m:
        diag_in = m["diagonal_size_meters"] / 0.0254
    else:
        diag_in = m.get("diagonal_size_inches")
    return geometry((W, H), distance_m=dist, fov_diagonal=m.get("fov_diagonal"), diagonal_size_inches=diag_in)


def geometry(resolution, distance_m=None, fov_diagonal=None, diagonal_size_inches=None):
    ar = resolution[0] / resolution[1]
    size_m = None
    if diagonal_size_inches is not None:
        h_mm = math.sqrt((diagonal_size_inches * 25.4) ** 2 / (1 + ar ** 2))
        size_m = (ar * h_mm / 1000, h_mm / 1000)
    if distance_m is None:
        distance_m = 3  # default for HMDs, fvvdp_display_model.py:406-408
    if fov_diagonal is not None:
        dist_px = math.sqrt(resolution[0] ** 2 + resolution[1] ** 2) / (2.0 * math.tan(math.radians(fov_diagonal * 0.5)))
        h_deg = math.degrees(math.atan(resolution[1] / 2 / dist_px)) * 2
        h_m = 2 * math.tan(math.radians(h_deg / 2)) * distance_m
        size_m = (h_m * ar, h_m)
    ppd = 1 / (2 * math.degrees(math.atan(0.5 * size_m[0] / resolution[0] / distance_m)))  # :436
    return dict(resolution=tuple(resolution), display_size_m=size_m, distance_m=distance_m, ppd_centre=ppd)


def eotf_forward(V, photo):
    """fvvdp_display_photo_eotf.forward :147-165 (srgb2lin :17-19, pq2lin :100-112);
    fvvdp_display_photo_absolute.forward :203-212 (kind 'absolute', keys L_min, L_max)."""
    V = V.astype(_F, copy=False)
    kind = photo["kind"]
    if kind == "absolute":
        return np.clip(V, _F(photo["L_min"]), _F(photo["L_max"]))
    Yp, Yb = _F(photo["Y_peak"]), _F(photo["Y_black"])
    if kind != "linear":
        V = np.clip(V, _F(0), _F(1))  # only applied (with a warning) when out of range; no-op otherwise
    if kind == "sRGB":
        lin = np.where(V > _F(0.04045), ((V + _F(0.055)) / _F(1.055)) ** _F(2.4), V / _F(12.92))
        return (Yp - Yb) * lin + Yb
    if kind == "gamma":
        return (Yp - Yb) * V ** _F(photo["gamma"]) + Yb
    if kind == "PQ":
        n, m_, c1, c2, c3 = 0.15930175781250000, 78.843750000000000, 0.83593750000000000, 18.851562500000000, 18.687500000000000
        t = V ** _F(1 / m_)
        L = _F(10000) * (np.maximum(t - _F(c1), _F(0)) / (_F(c2) - _F(c3) * t)) ** _F(1 / n)
        return np.clip(L, _F(0.005), Yp) + Yb
    if kind == "linear":
        return np.clip(V, _F(0.005), Yp) + Yb
    raise RuntimeError(f"Unknown EOTF '{kind}'")


def frame_luminance(frame, photo, rgb2y):
    """fvvdp_video_source_array._get_frame, video_source.py:180-208.  frame: (C,H,W) u8/u16/f32."""
    if frame.dtype == np.uint8:
        V = frame.astype(_F) / _F(255)
    elif frame.dtype == np.uint16:
        V = frame.astype(_F) / _F(65535)
    elif frame.dtype == np.float32:
        V = frame
    else:
        raise RuntimeError("Only uint8, uint16 and float32 is currently supported")
    L = eotf_forward(V, photo)
    if L.shape[0] == 3:
        L = L[0:1] * _F(rgb2y[0]) + L[1:2] * _F(rgb2y[1]) + L[2:3] * _F(rgb2y[2])
    return L[0]


# ----------------------------------------------------------------------------------------------
# temporal channels (fvvdp.py:228, 609-630, 258-300)
# ----------------------------------------------------------------------------------------------
def filter_len(fps):
    return int(np.ceil(250.0 / (1000.0 / fps)))  # fvvdp.py:228


def temporal_filters(fps, fl, sigma=0.5, beta=0.06):
    """get_temporal_filters fvvdp.py:609-630 -> F (2, fl) float32; F[0] sustained, F[1] transient."""
    t = np.linspace(0.0, fl / fps, fl, dtype=np.float64).astype(_F)
    F0 = np.exp(-((np.log(t + _F(1e-4)) - np.log(_F(beta))) ** _F(2.0)) / _F(2.0 * sigma ** 2.0)).astype(_F)
    F0 = F0 / np.sum(F0, dtype=_F)
    k2 = _F(0.062170507756932)
    F1 = np.concatenate([k2 * ((F0[1:] - F0[:-1]) / (t[1] - t[0])), np.zeros(1, _F)]).astype(_F)
    return np.stack([F0, F1], 0)


def window_indices(ff, N, fl, temp_padding):
    """Frame indices held by the sliding window (oldest first) when frame ff is scored
    (fvvdp.py:258-291).  The window is initialised at ff==0 and then slides by one frame."""
    if temp_padding == "replicate":
        first = [0] * fl
    elif temp_padding == "circular":
        first = [(N - 1 - fl + kk) % N for kk in range(fl)]
    elif temp_padding == "pingpong":
        pp = list(range(0, N)) + list(range(N - 2, 0, -1))
        idx = []
        while len(idx) < (fl - 1):
            idx = idx + pp
        first = idx[-(fl - 1):] + [0]
    else:
        raise RuntimeError(f'Unknown padding method "{temp_padding}"')
    seq = first + list(range(1, ff + 1))
    return seq[-fl:]


# ----------------------------------------------------------------------------------------------
# decimated Gaussian / contrast pyramid (fvvdp_lpyr_dec.py)
# ----------------------------------------------------------------------------------------------
_K = np.array([0.05, 0.25, 0.4, 0.25, 0.05], dtype=_F)  # get_kernels :176, kernel_a = 0.4


def pyramid_layout(W, H, ppd):
    """fvvdp_lpyr_dec.__init__ :15-49 -> (height, band_freqs[height+1])."""
    max_levels = int(np.floor(np.log2(min(H, W)))) - 1
    bands = np.concatenate([[1.0], np.power(2.0, -np.arange(0.0, 14.0)) * 0.3228], 0) * ppd / 2.0
    invalid = np.nonzero(bands <= 0.5)[0]
    max_band = max_levels if invalid.size == 0 else int(invalid[0])
    height = int(np.clip(max_band + 1, 0, max_levels))
    freqs = np.array([1.0] + [0.3228 * 2.0 ** (-f) for f in range(height)]) * ppd / 2.0
    return height, freqs


def _reduce_axis(x, axis, odd_rule):
    """5-tap, stride 2 along `axis`, zero-padded, plus the reference's edge terms (:188-205).
    odd_rule selects which last-sample correction is applied (the reference keys BOTH passes on the
    ROW count, :192 and :202 -- the second one is the documented quirk)."""
    x = np.moveaxis(x, axis, -1)
    n = x.shape[-1]
    m = (n + 1) // 2
    xp = np.zeros(x.shape[:-1] + (2 * m + 4,), _F)
    xp[..., 2:2 + n] = x
    out = np.zeros(x.shape[:-1] + (m,), _F)
    for k in range(5):
        out += _K[k] * xp[..., k:k + 2 * m:2]
    out[..., 0] += x[..., 0] * _K[1] + x[..., 1] * _K[0]
    if odd_rule:
        out[..., -1] += x[..., -1] * _K[3] + x[..., -2] * _K[4]
    else:
        out[..., -1] += x[..., -1] * _K[4]
    return np.moveaxis(out, -1, axis)


def gausspyr_reduce(x):
    """gausspyr_reduce :183-207.  x (..., H, W) float32."""
    H = x.shape[-2]
    y = _reduce_axis(x, -2, (H % 2) == 1)
    return _reduce_axis(y, -1, (H % 2) == 1)  # sic: row parity, fvvdp_lpyr_dec.py:202


def _expand_axis(x, S, axis):
    """interleave_zeros_and_pad :126-142 + conv with 2K (:228,233), as index arithmetic:
    z[m] = x[clamp(m/2-1)] for even m, 0 for odd m; out[i] = sum_k 2K[k] z[i+k]."""
    x = np.moveaxis(x, axis, -1)
    n = x.shape[-1]
    assert n == (S + 1) // 2
    i = np.arange(S)
    out = np.zeros(x.shape[:-1] + (S,), _F)
    for k in range(5):
        m = i + k
        sel = (m % 2) == 0
        j = np.clip(m[sel] // 2 - 1, 0, n - 1)
        out[..., sel] += (_F(2) * _K[k]) * x[..., j]
    return np.moveaxis(out, -1, axis)


def gausspyr_expand(x, sz):
    """gausspyr_expand :219-235 (rows first, then columns)."""
    return _expand_axis(_expand_axis(x, sz[0], -2), sz[1], -1)


def gaussian_pyramid(img, levels):
    res = [img]
    for _ in range(1, levels):
        res.append(gausspyr_reduce(res[-1]))
    return res


def contrast_pyramid(R4, height):
    """fvvdp_contrast_pyr.decompose :248-273.  R4 (C,H,W), channel 1 = reference sustained.
    Returns (bands[height+1] (last = base, Gaussian), L_bkg[height], gpyr)."""
    gpyr = gaussian_pyramid(R4, height + 1)
    bands, lbkg = [], []
    for i in range(height):
        ex = gausspyr_expand(gpyr[i + 1], gpyr[i].shape[-2:])
        layer = gpyr[i] - ex
        L = np.maximum(ex[1:2], _F(0.1))
        bands.append(np.minimum(layer / L, _F(1000.0)))
        lbkg.append(L[0])
    bands.append(gpyr[height])
    return bands, lbkg, gpyr


def band_mul(bb, n_bands_total):
    """get_band :57-63: first and last list element x1, others x2."""
    return _F(1.0) if (bb == 0 or bb == n_bands_total - 1) else _F(2.0)


def reconstruct(bands):
    """fvvdp_lpyr_dec.reconstruct :94-101."""
    img = bands[-1]
    for i in reversed(range(len(bands) - 1)):
        img = gausspyr_expand(img, bands[i].shape[-2:]) + bands[i]
    return img


# ----------------------------------------------------------------------------------------------
# CSF lookup (fvvdp.py:520-537, interp.py:11-59)
# ----------------------------------------------------------------------------------------------
def _interpolants(q, x):
    """get_interpolants_v1 interp.py:11-20."""
    imax = np.searchsorted(x, q, side="left")  # torch.bucketize(right=False)
    imax = np.minimum(imax, x.shape[0] - 1)
    imin = np.clip(imax - 1, 0, x.shape[0] - 1)
    frc = (q - x[imin]) / (x[imax] - x[imin] + _F(0.000001))
    frc = np.where(imax == imin, _F(0), frc)
    frc = np.where(frc < 0, _F(0), frc).astype(_F)
    return imin, imax, frc


def csf_sensitivity(rho, omega_idx, L_bkg, ecc):
    """cached_sensitivity fvvdp.py:520-537 (without the sensitivity_correction factor)."""
    lut = csf_lut()
    rho = np.broadcast_to(np.asarray(rho, _F), L_bkg.shape)
    ecc = np.broadcast_to(np.asarray(ecc, _F), L_bkg.shape)
    rq = np.log2(np.clip(rho, lut["rho"][0], lut["rho"][-1])).astype(_F).ravel()
    yq = np.log2(np.clip(L_bkg, lut["Y"][0], lut["Y"][-1])).astype(_F).ravel()
    eq = np.sqrt(np.clip(ecc, lut["ecc"][0], lut["ecc"][-1])).astype(_F).ravel()
    i0, i1, fi = _interpolants(rq, lut["rho_log"])
    j0, j1, fj = _interpolants(yq, lut["Y_log"])
    k0, k1, fk = _interpolants(eq, lut["ecc_sqrt"])
    v = lut["S_log"][omega_idx]
    one = _F(1)
    f = (((v[j0, i0, k0] * (one - fi) + v[j0, i1, k0] * fi) * (one - fj)
          + (v[j1, i0, k0] * (one - fi) + v[j1, i1, k0] * fi) * fj) * (one - fk)
         + ((v[j0, i0, k1] * (one - fi) + v[j0, i1, k1] * fi) * (one - fj)
            + (v[j1, i0, k1] * (one - fi) + v[j1, i1, k1] * fi) * fj) * fk)
    return np.power(_F(2.0), f.astype(_F)).reshape(L_bkg.shape)


# ----------------------------------------------------------------------------------------------
# foveation maps (fvvdp.py:416-442; fvvdp_display_model.py:475-526)
# ----------------------------------------------------------------------------------------------
def pix2view_direction(geo, res_wh, x_pix, y_pix):
    """fvvdp_display_model.py:498-510, degrees; x rightwards, y upwards."""
    xr = x_pix - _F(res_wh[0] / 2)
    yr = y_pix - _F(res_wh[1] / 2)
    x_m = xr * _F(geo["display_size_m"][0]) / _F(res_wh[0])
    y_m = -yr * _F(geo["display_size_m"][1]) / _F(res_wh[1])
    d = _F(geo["distance_m"])
    return np.rad2deg(np.arctan(x_m / d)).astype(_F), np.rad2deg(np.arctan(y_m / d)).astype(_F)


def resolution_magnification(geo, vx, vy):
    """get_ppd(view_dir)/get_ppd() :475-488, 512-526."""
    ppd_c = geo["ppd_centre"]
    va = np.minimum(np.sqrt(vx * vx + vy * vy), _F(89.9)).astype(_F)
    delta = (1 / ppd_c) / 2
    tan_delta = math.tan(math.radians(delta))
    tan_a = np.tan(np.deg2rad(va)).astype(_F)
    ppd = _F(ppd_c) * (np.tan(np.deg2rad(va + _F(delta))).astype(_F) - tan_a) / _F(tan_delta)
    return (ppd / _F(ppd_c)).astype(_F)


def foveation_maps(geo, band_hw, frame_hw, fixation_xy):
    """fvvdp.py:416-436.  A geometry dict may carry its own "pix2view_direction"(geo, res_wh, x, y) and
    "resolution_magnification"(geo, vx, vy) callables: the numpy twin of a fvvdp_display_geometry subclass
    (pytorch_examples/ex_custom_ppd.py:38-57)."""
    h, w = band_hw
    xv = np.linspace(0.5, w - 0.5, w, dtype=np.float64).astype(_F)
    yv = np.linspace(0.5, h - 0.5, h, dtype=np.float64).astype(_F)
    xx, yy = np.meshgrid(xv, yv, indexing="xy")
    p2v = geo.get("pix2view_direction", pix2view_direction)
    vx, vy = p2v(geo, (w, h), xx, yy)
    gx, gy = p2v(geo, (frame_hw[1], frame_hw[0]), _F(fixation_xy[0]) + _F(0.5), _F(fixation_xy[1]) + _F(0.5))
    ecc = np.sqrt((vx - gx) ** 2 + (vy - gy) ** 2).astype(_F)
    return ecc, geo.get("resolution_magnification", resolution_magnification)(geo, vx, vy)


# ----------------------------------------------------------------------------------------------
# masking, pooling (fvvdp.py:574-607, 337-357)
# ----------------------------------------------------------------------------------------------
def masking(T, R, S, cc, p):
    """apply_masking_model fvvdp.py:574-596 with N = 1/S."""
    N = _F(1) / S
    q = _F(p["mask_q_sust"] if cc == 0 else p["mask_q_trans"])
    Tn, Rn = T / N, R / N
    M = np.minimum(np.abs(Tn), np.abs(Rn)) * np.power(_F(10.0), _F(p["mask_c"]))
    D = np.power(np.abs(Tn - Rn), _F(p["mask_p"])) / (_F(1) + np.power(M, q))
    return np.minimum(D, _F(1e4))


def lp_norm_spatial(D, beta):
    """lp_norm(D.flatten(), beta, 0, True) fvvdp.py:598-607 (accumulated in float64 like torch.norm's
    pairwise float32 sum to within ~1e-7 relative)."""
    s = np.sum(np.power(D.astype(np.float64), beta))
    return _F((s ** (1.0 / beta)) / (float(D.size) ** (1.0 / beta)))


def pool_to_jod(Q_per_ch, p, is_video=True):
    """do_pooling_and_jods fvvdp.py:337-357.  Q_per_ch (bands, 2, N)."""
    Q = Q_per_ch.astype(_F)
    if is_video or Q.shape[1] == 2:
        w = np.array([1.0, p["w_transient"]], _F)[None, :, None]
        Q = Q * w

    def lp(x, b, dim, normalize):
        n = x.shape[dim] if normalize else 1.0
        return (np.sum(np.abs(x).astype(np.float64) ** b, axis=dim, keepdims=True) ** (1.0 / b) / (float(n) ** (1.0 / b))).astype(_F)

    Q_sc = lp(Q, p["beta_sch"], 0, False)
    Q_tc = lp(Q_sc, p["beta_tch"], 1, False)
    Qv = float(lp(Q_tc, p["beta_t"], 2, True).squeeze())
    beta_jod = 10.0 ** p["log_jod_exp"]
    sign = -1.0 if p["jod_a"] < 0 else 1.0
    return sign * ((abs(p["jod_a"]) ** (1.0 / beta_jod)) * Qv) ** beta_jod + 10.0


# ----------------------------------------------------------------------------------------------
# heat-map visualisation (visualize_diff_map.py:9-107, interp.py:72-80)
# ----------------------------------------------------------------------------------------------
_COLOR_MAPS = {  # visualize_diff_map.py:66-82
    "threshold": ([[0.2, 0.2, 1.0], [0.2, 1.0, 1.0], [0.2, 1.0, 0.2], [1.0, 1.0, 0.2], [1.0, 0.2, 0.2]], [0.00, 0.25, 0.50, 0.75, 1.00]),
    "supra-threshold": ([[0.2, 1.0, 1.0], [1.0, 1.0, 1.0], [1.0, 1.0, 0.2]], [0.0, 0.5, 1.0]),
}


def linspace_f32(lo, hi, n):
    """torch.linspace in float32: start + step*i in the first half, end - step*(n-1-i) in the second."""
    lo, hi = _F(lo), _F(hi)
    step = _F((hi - lo) / _F(n - 1))
    i = np.arange(n)
    return np.where(i < n // 2, lo + step * i.astype(_F), hi - step * (n - 1 - i).astype(_F)).astype(_F)


def interp1(x, v, q):
    """interp.py:72-80."""
    imin, imax, frc = _interpolants(q.ravel().astype(_F), x)
    return (v[imin] * (_F(1) - frc) + v[imax] * frc).astype(_F).reshape(q.shape)


def vis_tonemap(b, dr):
    """visualize_diff_map.py:26-50: histogram-equalising tone curve with exponent 1/3 over 1024 bins."""
    dr = _F(dr)
    b_min, b_max = b.min(), b.max()
    if b_max - b_min < dr:
        return ((b - b_min) / (b_max - b_min + _F(1e-3)) * dr + (_F(1) - dr) / _F(2)).astype(_F)
    b_scale = linspace_f32(b_min, b_max, 1024)
    # torch.histc: bin = (int)((x - min) / (max - min) * bins), the maximum goes into the last bin
    pos = ((b.ravel() - b_min) / (b_max - b_min) * _F(1024)).astype(np.int64)
    pos = np.minimum(pos, 1023)
    b_p = np.bincount(pos, minlength=1024).astype(_F)
    b_p = b_p / np.sum(b_p, dtype=_F)
    pw = np.power(b_p, _F(1.0 / 3.0)).astype(_F)
    dy = pw / np.sum(pw, dtype=_F)
    v = (np.cumsum(dy, dtype=_F) * dr + (_F(1) - dr) / _F(2)).astype(_F)
    return interp1(b_scale, v, b)


def visualize_diff_map(diff_map, context, colormap_type):
    """visualize_diff_map.py:58-107 for a single-channel context image.  diff_map, context (H,W) float32
    -> (3,H,W) float32 in [0,1]."""
    d = np.clip(diff_map.astype(_F), _F(0), _F(1))
    y = context.astype(_F)
    clampval = y[y > 0].min()
    tmo = vis_tonemap(np.log(np.maximum(y, clampval)).astype(_F), 0.6)
    cm, cm_in = _COLOR_MAPS[colormap_type]
    cm = np.array(cm, _F)
    cm_in = np.array(cm_in, _F)
    cm_l = cm[:, 0:1] * _F(0.212656) + cm[:, 1:2] * _F(0.715158) + cm[:, 2:3] * _F(0.072186)
    cm_ch = cm / (cm_l + _F(0.0001))
    out = np.stack([interp1(cm_in, cm_ch[:, c], d) for c in range(3)], 0)
    return np.clip(out * tmo[None], _F(0), _F(1)).astype(_F)


# ----------------------------------------------------------------------------------------------
# planar Y'CbCr frames (video_source_yuv.py:157-228, video_source_file.py:219-276)
# ----------------------------------------------------------------------------------------------
_YCBCR2RGB = {"2020": [[1, 0, 1.47460], [1, -0.16455, -0.57135], [1, 1.88140, 0]],
              "709": [[1, 0, 1.402], [1, -0.344136, -0.714136], [1, 1.772, 0]]}


def _upsample2_bilinear(c):
    """torch.nn.functional.interpolate(scale_factor=2, mode='bilinear') (align_corners=False) of a (h,w) plane:
    source coordinate max((dst + 0.5) / 2 - 0.5, 0), second tap clamped to the last sample."""
    h, w = c.shape

    def taps(n):
        src = np.maximum((np.arange(2 * n, dtype=_F) + _F(0.5)) * _F(0.5) - _F(0.5), _F(0))
        i0 = src.astype(np.int64)
        return i0, np.minimum(i0 + 1, n - 1), (src - i0.astype(_F)).astype(_F)

    y0, y1, ly = taps(h)
    x0, x1, lx = taps(w)
    ly, lx = ly[:, None], lx[None, :]
    top = (_F(1) - lx) * c[y0][:, x0] + lx * c[y0][:, x1]
    bot = (_F(1) - lx) * c[y1][:, x0] + lx * c[y1][:, x1]
    return ((_F(1) - ly) * top + ly * bot).astype(_F)


def yuv_frame_rgb(Y, u, v, bit_depth, chroma_ss, color_space):
    """One planar frame (integer planes) -> display-encoded RGB (H,W,3) float32 in [0,1]:
    _fixed2float_upscale :198-228 and get_frame_rgb_tensor :157-182."""
    sc = _F(2 ** (bit_depth - 8))
    Yf = np.clip(_F(1) / (sc * _F(219)) * Y.astype(_F) - _F(16 / 219), _F(0), _F(1))
    planes = []
    for c in (u, v):
        cf = np.clip(_F(1) / (sc * _F(224)) * c.astype(_F) - _F(128 / 224), _F(-0.5), _F(0.5)).astype(_F)
        planes.append(_upsample2_bilinear(cf) if chroma_ss == "420" else cf)
    yuv = np.stack([Yf, planes[0], planes[1]], -1).astype(_F)
    M = np.array(_YCBCR2RGB["2020" if color_space == "2020" else "709"], _F)
    return np.clip(yuv @ M.T, _F(0), _F(1)).astype(_F)


def resize_rgb(rgb, out_w, out_h, mode):
    """Full-screen resize of a display-encoded (H,W,3) frame (fvvdp_video_source_yuv_file._get_frame, video_source_yuv.py:293-297):
    torch.nn.functional.interpolate(size=(out_h, out_w), mode=mode), align_corners unset, then clip to [0,1].  Tap positions and
    weights restate ATen's upsample kernels: scale = in / out (float32); nearest: min(floor(dst * scale), in - 1); bilinear:
    src = max(scale (dst + 0.5) - 0.5, 0), second tap clamped; bicubic: src = scale (dst + 0.5) - 0.5, A = -0.75, taps clamped to
    the frame; area: mean over [floor(i in / out), ceil((i + 1) in / out))."""
    rgb = np.asarray(rgb, _F)
    H, W = rgb.shape[:2]

    def lin_taps(n_out, n_in):
        src = np.maximum(_F(n_in) / _F(n_out) * (np.arange(n_out, dtype=_F) + _F(0.5)) - _F(0.5), _F(0))
        i0 = src.astype(np.int64)
        return i0, i0 + (i0 < n_in - 1), (src - i0.astype(_F)).astype(_F)

    def cubic_taps(n_out, n_in):
        src = _F(n_in) / _F(n_out) * (np.arange(n_out, dtype=_F) + _F(0.5)) - _F(0.5)
        b = np.floor(src)
        t = (src - b).astype(_F)
        A = _F(-0.75)
        c1 = lambda x: ((A + _F(2)) * x - (A + _F(3))) * x * x + _F(1)
        c2 = lambda x: ((A * x - _F(5) * A) * x + _F(8) * A) * x - _F(4) * A
        w = np.stack([c2(t + _F(1)), c1(t), c1(_F(1) - t), c2(_F(2) - t)], 1).astype(_F)
        idx = np.clip(b.astype(np.int64)[:, None] + np.arange(-1, 3)[None, :], 0, n_in - 1)
        return idx, w

    if mode == "nearest":
        iy = np.minimum(np.floor(np.arange(out_h, dtype=_F) * (_F(H) / _F(out_h))).astype(np.int64), H - 1)
        ix = np.minimum(np.floor(np.arange(out_w, dtype=_F) * (_F(W) / _F(out_w))).astype(np.int64), W - 1)
        out = rgb[iy][:, ix]
    elif mode == "bilinear":
        y0, y1, ly = lin_taps(out_h, H)
        x0, x1, lx = lin_taps(out_w, W)
        ly, lx = ly[:, None, None], lx[None, :, None]
        top = (_F(1) - lx) * rgb[y0][:, x0] + lx * rgb[y0][:, x1]
        bot = (_F(1) - lx) * rgb[y1][:, x0] + lx * rgb[y1][:, x1]
        out = (_F(1) - ly) * top + ly * bot
    elif mode == "bicubic":
        iy, wy = cubic_taps(out_h, H)
        ix, wx = cubic_taps(out_w, W)
        rows = np.zeros((H, out_w, 3), _F)
        for k in range(4):
            rows += rgb[:, ix[:, k]] * wx[None, :, k, None]
        out = np.zeros((out_h, out_w, 3), _F)
        for k in range(4):
            out += rows[iy[:, k]] * wy[:, k, None, None]
    elif mode == "area":
        ys = [((i * H) // out_h, -((-(i + 1) * H) // out_h)) for i in range(out_h)]
        xs = [((i * W) // out_w, -((-(i + 1) * W) // out_w)) for i in range(out_w)]
        cols = np.stack([rgb[:, a:b].sum(1, dtype=_F) for a, b in xs], 1)                      # (H, out_w, 3)
        out = np.stack([cols[a:b].sum(0, dtype=_F) for a, b in ys], 0)
        out = out / (np.array([b - a for a, b in ys], _F)[:, None, None] * np.array([b - a for a, b in xs], _F)[None, :, None])
    else:
        raise ValueError(f"unknown resize mode {mode}")
    return np.clip(out, _F(0), _F(1)).astype(_F)


# ----------------------------------------------------------------------------------------------
# PU21-PSNR (pupsnr.py:52-79, utils.py:157-202)
# ----------------------------------------------------------------------------------------------
_PU21 = [234.0235618, 216.9339286, 0.0001091864237, 0.893206924, 0.06733984121, 1.444718567, 567.6315065]  # 'banding_glare'


def pu_encode(Y, L_min=0.005, L_max=10000.0):
    p = _PU21
    Y = np.clip(Y.astype(_F), _F(L_min), _F(L_max))
    Yp = Y ** _F(p[3])
    return (_F(p[6]) * (((_F(p[0]) + _F(p[1]) * Yp) / (_F(1) + _F(p[2]) * Yp)) ** _F(p[4]) - _F(p[5]))).astype(_F)


def pu_psnr(test, ref, dim_order="BCFHW", display_name="standard_4k", photometry=None, color_space="sRGB"):
    """pu_psnr.predict_video_source: mean over the frames of 20 log10(peak / sqrt(mean((PU(T) - PU(R))^2)))."""
    p = _PU21
    L_max = 10000.0
    peak = p[6] * (((p[0] + p[1] * L_max ** p[3]) / (1 + p[2] * L_max ** p[3])) ** p[4] - p[5])
    photo = photometry if photometry is not None else photometry_from_preset(display_name)
    rgb2y = metric_data()["rgb2y"][color_space]
    tv, rv = to_bcfhw(np.asarray(test), dim_order), to_bcfhw(np.asarray(ref), dim_order)
    N = tv.shape[2]
    total = 0.0
    for ff in range(N):
        d = pu_encode(frame_luminance(tv[0, :, ff], photo, rgb2y)) - pu_encode(frame_luminance(rv[0, :, ff], photo, rgb2y))
        mse = np.mean(d.astype(np.float64) ** 2)
        total += 20.0 * math.log10(peak / math.sqrt(mse)) / N
    return total


# ----------------------------------------------------------------------------------------------
# the metric (fvvdp.py:190-334, 359-478)
# ----------------------------------------------------------------------------------------------
def to_bcfhw(a, dim_order):
    """reshuffle_dims video_source.py:43-69 -> (B,C,F,H,W) view."""
    dim_order = dim_order.upper()
    out = "BCFHW"
    inter = [c for c in out if c in dim_order]
    a = np.transpose(a, [dim_order.index(c) for c in inter])
    shape = [a.shape[inter.index(c)] if c in inter else 1 for c in out]
    return a.reshape(shape)


def score_frame(R4, height, freqs, p, is_image, foveated=False, geo=None, frame_hw=None, fixation_xy=None, taps=None, want_heatmap=False):
    """process_block_of_frames fvvdp.py:359-478 for one frame.  R4 (4,H,W) [T_s,R_s,T_t,R_t]
    (image: (2,H,W))."""
    bands, lbkg, gpyr = contrast_pyramid(R4, height)
    nb = height + 1
    temp_ch = 1 if is_image else 2
    Q = np.zeros((height, 2), _F)
    sens_mul = _F(10.0 ** (p["sensitivity_correction"] / 20.0))
    w_ch = [_F(1.0), _F(p["w_transient"])]
    hm_bands = [None] * height
    for cc in range(temp_ch):
        for bb in range(height):
            m = band_mul(bb, nb)
            T_f = bands[bb][cc * 2 + 0] * m
            R_f = bands[bb][cc * 2 + 1] * m
            L = lbkg[bb]
            if foveated:
                ecc, res_mag = foveation_maps(geo, T_f.shape, frame_hw, fixation_xy)
            else:
                ecc, res_mag = np.zeros(T_f.shape, _F), np.ones(T_f.shape, _F)
            rho = _F(freqs[bb]) * res_mag
            S = csf_sensitivity(rho, cc, L, ecc) * sens_mul
            D = masking(T_f, R_f, S, cc, p)
            Q[bb, cc] = lp_norm_spatial(D, p["beta"])
            if want_heatmap:
                hm_bands[bb] = (D if cc == 0 else hm_bands[bb] * m + w_ch[cc] * D) / m  # set_band/get_band :57-71
            if taps is not None:
                taps.setdefault("T_f", {})[(bb, cc)] = T_f
                taps.setdefault("R_f", {})[(bb, cc)] = R_f
                taps.setdefault("L_bkg", {})[bb] = L
                taps.setdefault("S", {})[(bb, cc)] = S
                taps.setdefault("D", {})[(bb, cc)] = D
    if taps is not None:
        taps["gpyr"] = gpyr
    dmap = None
    if want_heatmap:
        beta_jod = 10.0 ** p["log_jod_exp"]
        rec = reconstruct(hm_bands + [np.zeros(gpyr[height].shape[-2:], _F)])
        dmap = (np.power(rec, _F(beta_jod)) * _F(abs(p["jod_a"]))).astype(_F)
    return Q, dmap


def predict(test, ref, dim_order="BCFHW", frames_per_second=0, display_name="standard_4k", photometry=None,
            geometry_=None, color_space="sRGB", foveated=False, fixation_point=None, temp_padding="replicate",
            heatmap=None, frames=None, tap_frame=None, lum_cache=None, frame_times=None):
    """fvvdp.predict / predict_video_source, fvvdp.py:181-334.

    frames: optional iterable of frame indices to score (the others are skipped; used by the bounded
    CPU-baseline sample and the sharding tests).  Returns (jod, stats); stats['taps'] holds the
    intermediate tensors of frame `tap_frame` when requested.  lum_cache: optional dict {(stream, frame): luminance}
    of already converted frames (bench.py pre-fills the temporal window so that the timed region holds the
    steady-state per-frame work only); frame_times: optional list that receives the seconds spent per scored frame."""
    import time
    md = metric_data()
    p = md["parameters"]
    photo = photometry if photometry is not None else photometry_from_preset(display_name)
    geo = geometry_ if geometry_ is not None else geometry_from_preset(display_name)
    rgb2y = md["rgb2y"][color_space]
    tv, rv = to_bcfhw(np.asarray(test), dim_order), to_bcfhw(np.asarray(ref), dim_order)
    assert tv.shape == rv.shape
    _, C, N, H, W = tv.shape
    is_image = N == 1
    height, freqs = pyramid_layout(W, H, geo["ppd_centre"])
    if fixation_point is None:
        fixation_point = np.array([W // 2, H // 2])
    fixation_point = np.asarray(fixation_point)
    score = list(range(N)) if frames is None else list(frames)
    Q_per_ch = np.zeros((height, 2, N), _F)
    want_hm = heatmap not in (None, "none")
    hm = np.zeros((1, 1 if heatmap == "raw" else 3, N, H, W), np.float16) if want_hm else None
    taps = None
    lum_cache = {} if lum_cache is None else lum_cache

    def lum(which, idx):
        key = (which, idx)
        if key not in lum_cache:
            src = tv if which == 0 else rv
            lum_cache[key] = frame_luminance(src[0, :, idx], photo, rgb2y)
        return lum_cache[key]

    if not is_image:
        fl = filter_len(frames_per_second)
        F = temporal_filters(frames_per_second, fl, p["sustained_sigma"], p["sustained_beta"])
    for ff in score:
        t_start = time.perf_counter()
        if is_image:
            R4 = np.stack([lum(0, 0), lum(1, 0)], 0)
        else:
            win = window_indices(ff, N, fl, temp_padding)
            R4 = np.zeros((4, H, W), _F)
            for cc in range(2):
                w = F[cc][::-1]  # corr_filter = F.flip(0), fvvdp.py:298
                for s in range(2):
                    acc = np.zeros((H, W), _F)
                    for k in range(fl):
                        acc += lum(s, win[k]) * w[k]
                    R4[cc * 2 + s] = acc
            keep = set(window_indices(min(ff + 1, N - 1), N, fl, temp_padding))
            for key in [k for k in lum_cache if k[1] not in keep]:
                del lum_cache[key]
        fx = fixation_point[ff] if fixation_point.ndim == 2 else fixation_point
        t = {} if (tap_frame is not None and ff == tap_frame) else None
        Q, dmap = score_frame(R4, height, freqs, p, is_image, foveated, geo, (H, W), fx, t, want_hm)
        if t is not None:
            t["R"] = R4
            taps = t
        Q_per_ch[:, :, ff] = Q
        if want_hm:
            if heatmap == "raw":
                hm[0, 0, ff] = dmap.astype(np.float16)
            else:  # the context image is R[:,0], the sustained channel of the TEST stream (fvvdp.py:475)
                hm[0, :, ff] = visualize_diff_map(dmap, R4[0], heatmap).astype(np.float16)
        if frame_times is not None:
            frame_times.append(time.perf_counter() - t_start)
    sel = Q_per_ch if frames is None else Q_per_ch[:, :, score]
    jod = pool_to_jod(sel, p)
    stats = dict(Q_per_ch=Q_per_ch, rho_band=freqs, frames_per_second=frames_per_second, width=W, height=H, N_frames=N)
    if want_hm:
        stats["heatmap"] = hm
    if taps is not None:
        stats["taps"] = taps
    return float(jod), stats
