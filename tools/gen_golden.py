#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (pyfvvdp, torch CPU fp32) from
/root/reference in the build container.  The fixtures pin oracle/fvvdp_oracle.py (and through it the
CUDA path).  Inputs are analytic / seeded so only outputs (and small inputs) are stored.

    python tools/gen_golden.py            # rewrites every fixture
    python tools/gen_golden.py widen      # only the fixtures of the widened surface (heat-map colour maps, custom geometry)
    python tools/gen_golden.py yuv        # only the raw .yuv video-source fixtures
    python tools/gen_golden.py yuvresize  # only the full-screen-resize fixtures of .yuv clips
    python tools/gen_golden.py round2     # only the fixtures added in round 2 (GOG, 120 fps, colour-space mismatch, the 64-frame 4K bench clip)

Large tap tensors are stored as strided sub-samples ([::SY, ::SX]) to keep the fixtures small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from _refimport import import_reference  # noqa: E402

pyfvvdp = import_reference()
import torch  # noqa: E402
from pyfvvdp.fvvdp_lpyr_dec import fvvdp_lpyr_dec, fvvdp_contrast_pyr  # noqa: E402
from pyfvvdp.fvvdp_display_model import (fvvdp_display_photo_eotf, fvvdp_display_photo_absolute,  # noqa: E402
                                         fvvdp_display_photometry, fvvdp_display_geometry)

from fovvideovdp_b200.synthetic import synth_pair_numpy  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CPU = torch.device("cpu")
SY, SX = 5, 7


def sub(a):
    a = np.asarray(a)
    if a.size <= 40000:
        return a
    return a[..., ::SY, ::SX]


def run_with_taps(fv, test, ref, tap_frame, **kw):
    """predict() while recording R, contrast bands (x band_mul), L_bkg, S, D of one frame."""
    taps = {}
    orig_block = fv.process_block_of_frames
    orig_sens = fv.cached_sensitivity
    orig_mask = fv.apply_masking_model
    state = {"on": False, "s": [], "d": []}

    def block(ff, R, *a, **k):
        state["on"] = ff == tap_frame
        if state["on"]:
            taps["R"] = R[0, :, 0].numpy().copy()
        out = orig_block(ff, R, *a, **k)
        state["on"] = False
        return out

    def sens(rho, omega, L_bkg, ecc, sigma):
        S = orig_sens(rho, omega, L_bkg, ecc, sigma)
        if state["on"]:
            state["s"].append((L_bkg.numpy().copy(), S.numpy().copy()))
        return S

    def mask(T, R, N, cc):
        D = orig_mask(T, R, N, cc)
        if state["on"]:
            state["d"].append((cc, T.numpy().copy(), R.numpy().copy(), D.numpy().copy()))
        return D

    fv.process_block_of_frames = block
    fv.cached_sensitivity = sens
    fv.apply_masking_model = mask
    try:
        q, stats = fv.predict(torch.tensor(test), torch.tensor(ref), **kw)
    finally:
        fv.process_block_of_frames, fv.cached_sensitivity, fv.apply_masking_model = orig_block, orig_sens, orig_mask
    nb = fv.lpyr.height
    out = {"R": sub(taps["R"])}
    sens_mul = np.float32(10.0 ** (fv.sensitivity_correction / 20.0))
    for i, (cc, T, R, D) in enumerate(state["d"]):
        bb = i % nb
        L, S = state["s"][i]
        out[f"T_f_{bb}_{cc}"] = sub(T)
        out[f"R_f_{bb}_{cc}"] = sub(R)
        out[f"D_{bb}_{cc}"] = sub(D)
        out[f"S_{bb}_{cc}"] = sub(np.broadcast_to(S.reshape(S.shape[-2:]) * sens_mul, T.shape))
        out[f"L_bkg_{bb}"] = sub(L.reshape(L.shape[-2:]))
    return float(q), stats, out


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB")


def gen_metric_cases():
    test, ref = synth_pair_numpy(12, 270, 480)
    # K1: FHD video, replicate padding, taps of frame 5
    fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU)
    q, st, taps = run_with_taps(fv, test, ref, 5, dim_order="BCFHW", frames_per_second=30)
    save("video_fhd_replicate", jod=q, Q_per_ch=st["Q_per_ch"], rho_band=st["rho_band"], tap_frame=5, sub=[SY, SX], **taps)
    # K4/K5: padding modes
    for pad in ("pingpong", "circular"):
        fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU, temp_padding=pad)
        q, st = fv.predict(torch.tensor(test), torch.tensor(ref), dim_order="BCFHW", frames_per_second=30)
        save(f"video_fhd_{pad}", jod=float(q), Q_per_ch=st["Q_per_ch"])
    # short clip (N < filter_len) with every padding mode, 25 fps (fl=7) and 60 fps (fl=15)
    for fps in (25, 60):
        for pad in ("replicate", "pingpong", "circular"):
            fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU, temp_padding=pad)
            q, st = fv.predict(torch.tensor(test[:, :, :5]), torch.tensor(ref[:, :, :5]), dim_order="BCFHW", frames_per_second=fps)
            save(f"video_short_{fps}fps_{pad}", jod=float(q), Q_per_ch=st["Q_per_ch"])
    # K2: image
    fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU)
    q, st, taps = run_with_taps(fv, test[0, :, 0:1], ref[0, :, 0:1], 0, dim_order="CFHW")
    save("image_fhd", jod=q, Q_per_ch=st["Q_per_ch"], **taps)
    # K3: foveated HDR PQ, moving gaze
    gaze = np.stack([np.linspace(0, 479, 12), np.linspace(0, 269, 12)], 1).astype(np.float32)
    fv = pyfvvdp.fvvdp(display_name="standard_hdr_pq", device=CPU, foveated=True)
    tq, rq = 0.1 + 0.65 * test, 0.1 + 0.65 * ref
    q, st, taps = run_with_taps(fv, tq, rq, 5, dim_order="BCFHW", frames_per_second=30, fixation_point=gaze)
    save("video_hdrpq_foveated", jod=q, Q_per_ch=st["Q_per_ch"], gaze=gaze, tap_frame=5, **taps)
    # foveated, fixed gaze given as (2,), on the wide-FOV HMD preset (strong resolution magnification)
    fv = pyfvvdp.fvvdp(display_name="standard_hmd", device=CPU, foveated=True)
    q, st = fv.predict(torch.tensor(test[:, :, :4]), torch.tensor(ref[:, :, :4]), dim_order="BCFHW", frames_per_second=30,
                       fixation_point=np.array([100.0, 50.0], dtype=np.float32))
    save("video_hmd_foveated_fixed", jod=float(q), Q_per_ch=st["Q_per_ch"])
    # K6: heatmap raw
    fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU, heatmap="raw")
    q, st = fv.predict(torch.tensor(test[:, :, :4]), torch.tensor(ref[:, :, :4]), dim_order="BCFHW", frames_per_second=30)
    hm = st["heatmap"].float().numpy()
    save("video_fhd_heatmap_raw", jod=float(q), Q_per_ch=st["Q_per_ch"], heatmap_sub=hm[0, 0, :, ::SY, ::SX], hm_mean=hm.mean(), hm_max=hm.max())
    # odd sizes / parity quirk / other dtypes and dim orders
    for (H, W) in ((135, 240), (136, 241), (67, 97), (64, 64)):
        t2, r2 = synth_pair_numpy(3, H, W)
        fv = pyfvvdp.fvvdp(display_name="standard_4k", device=CPU)
        q, st, taps = run_with_taps(fv, t2, r2, 2, dim_order="BCFHW", frames_per_second=24)
        save(f"video_4k_{H}x{W}", jod=q, Q_per_ch=st["Q_per_ch"], tap_frame=2, **taps)
    # uint8 RGB, FHWC
    t3, r3 = synth_pair_numpy(3, 72, 3 * 50)
    t8 = np.round(t3[0, 0].reshape(3, 72, 50, 3) * 255).astype(np.uint8)
    r8 = np.round(r3[0, 0].reshape(3, 72, 50, 3) * 255).astype(np.uint8)
    fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU)
    q, st = fv.predict(t8, r8, dim_order="FHWC", frames_per_second=30)
    save("video_u8_rgb_fhwc", jod=float(q), Q_per_ch=st["Q_per_ch"], test=t8, ref=r8)
    # uint16 RGB image, HWC, BT.2020 + gamma EOTF custom photometry
    t16 = np.round(t3[0, 0].reshape(3, 72, 50, 3)[0] * 65535).astype(np.uint16)
    r16 = np.round(r3[0, 0].reshape(3, 72, 50, 3)[0] * 65535).astype(np.uint16)
    ph = fvvdp_display_photo_eotf(400, contrast=2000, EOTF="gamma", gamma=2.4, E_ambient=100)
    fv = pyfvvdp.fvvdp(display_name="standard_4k", display_photometry=ph, color_space="BT.2020", device=CPU)
    q, st = fv.predict(t16, r16, dim_order="HWC")
    save("image_u16_rgb_gamma_bt2020", jod=float(q), Q_per_ch=st["Q_per_ch"], test=t16, ref=r16)
    # absolute photometry (linear cd/m^2 input), luminance-only video
    ph = fvvdp_display_photo_absolute(L_max=4000, L_min=0.01)
    fv = pyfvvdp.fvvdp(display_name="standard_4k", display_photometry=ph, device=CPU)
    ta, ra = (t2 * 300 + 0.001).astype(np.float32), (r2 * 300 + 0.001).astype(np.float32)
    q, st = fv.predict(ta, ra, dim_order="BCFHW", frames_per_second=30)
    save("video_absolute", jod=float(q), Q_per_ch=st["Q_per_ch"])
    # linear EOTF preset
    fv = pyfvvdp.fvvdp(display_name="standard_hdr_linear", device=CPU)
    q, st = fv.predict(ta, ra, dim_order="BCFHW", frames_per_second=30)
    save("video_hdr_linear", jod=float(q), Q_per_ch=st["Q_per_ch"])


def gen_unit_cases():
    rng = np.random.default_rng(1234)
    pyr = fvvdp_lpyr_dec(64, 64, 30.0, CPU)
    arrs = {}
    for i, (H, W) in enumerate([(135, 240), (136, 241), (136, 240), (135, 241), (17, 30), (33, 47), (5, 7), (4, 4), (9, 8), (68, 120)]):
        x = (rng.random((2, 1, H, W), dtype=np.float32) * 100).astype(np.float32)
        y = pyr.gausspyr_reduce(torch.tensor(x)).numpy()
        e = pyr.gausspyr_expand(torch.tensor(y), [H, W]).numpy()
        arrs[f"x{i}"], arrs[f"red{i}"], arrs[f"exp{i}"] = x, y, e
    save("unit_pyramid", n=10, **arrs)

    # contrast pyramid on a 4-channel stack
    cp = fvvdp_contrast_pyr(97, 67, 37.84, CPU)
    x = (rng.random((4, 1, 67, 97), dtype=np.float32) * 200 + 0.05).astype(np.float32)
    bands, lbkg = cp.decompose(torch.tensor(x))
    d = {"x": x, "height": cp.height, "freqs": cp.get_freqs()}
    for i, b in enumerate(bands):
        d[f"band{i}"] = b.numpy()
    for i, b in enumerate(lbkg):
        d[f"lbkg{i}"] = b.numpy()
    save("unit_contrast_pyr", **d)

    # pyramid layout
    rows = []
    for (W, H, ppd) in [(1920, 1080, 37.84), (3840, 2160, 75.40), (480, 270, 37.84), (1024, 683, 75.4), (64, 64, 75.4),
                        (97, 67, 20.0), (1440, 1600, 11.9), (240, 135, 60.0), (32, 16, 5.0), (8000, 4000, 200.0)]:
        p = fvvdp_lpyr_dec(W, H, ppd, CPU)
        rows.append([W, H, ppd, p.height] + list(p.get_freqs()) + [0.0] * (16 - len(p.get_freqs())))
    save("unit_pyr_layout", rows=np.array(rows, dtype=np.float64))

    # CSF interpolation + masking
    fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU)
    n = 4096
    rho = np.exp(rng.uniform(np.log(0.03), np.log(90), n)).astype(np.float32)
    L = np.exp(rng.uniform(np.log(5e-4), np.log(3e4), n)).astype(np.float32)
    ecc = rng.uniform(0, 130, n).astype(np.float32)
    ecc[:64] = 0.0
    lut = fv.csf_cache[fv.get_cache_key(0, fv.csf_sigma, fv.k_cm)]["lut"]
    rho[64:96] = lut["rho"].numpy()  # exact grid hits
    L[96:128] = lut["Y"].numpy()
    S0 = fv.cached_sensitivity(torch.tensor(rho), fv.omega[0], torch.tensor(L), torch.tensor(ecc), fv.csf_sigma).numpy()
    S5 = fv.cached_sensitivity(torch.tensor(rho), fv.omega[1], torch.tensor(L), torch.tensor(ecc), fv.csf_sigma).numpy()
    T = (rng.standard_normal(n) * np.exp(rng.uniform(-6, 3, n))).astype(np.float32)
    R = (T + rng.standard_normal(n) * np.exp(rng.uniform(-8, 1, n))).astype(np.float32)
    T[:8], R[:8] = 0.0, 0.0
    R[8:16] = T[8:16]
    Smul = (S0 * np.float32(10.0 ** (fv.sensitivity_correction / 20.0))).astype(np.float32)
    D0 = fv.apply_masking_model(torch.tensor(T), torch.tensor(R), torch.reciprocal(torch.tensor(Smul)), 0).numpy()
    D1 = fv.apply_masking_model(torch.tensor(T), torch.tensor(R), torch.reciprocal(torch.tensor(Smul)), 1).numpy()
    save("unit_csf_masking", rho=rho, L=L, ecc=ecc, S0=S0, S5=S5, T=T, R=R, Smul=Smul, D0=D0, D1=D1)

    # temporal filters
    d = {}
    for fps in (24, 25, 30, 50, 60, 120, 12.5):
        fv.filter_len = int(np.ceil(250.0 / (1000.0 / fps)))
        F, _ = fv.get_temporal_filters(fps)
        d[f"F_{fps}"] = F.numpy()
    save("unit_temporal_filters", **d)

    # EOTFs
    V = np.linspace(-0.1, 1.1, 241, dtype=np.float32)
    d = {"V": V}
    import logging
    logging.disable(logging.WARNING)
    for kind in ("sRGB", "gamma", "PQ", "linear"):
        ph = fvvdp_display_photo_eotf(1500 if kind in ("PQ", "linear") else 200, contrast=1000, EOTF=kind, gamma=2.2, E_ambient=250)
        Vin = V * 2000 if kind == "linear" else V
        d[kind] = ph.forward(torch.tensor(Vin)).numpy()
        d[kind + "_black"] = ph.get_black_level()
    d["absolute"] = fvvdp_display_photo_absolute(1000, 0.01).forward(torch.tensor(V * 2000)).numpy()
    logging.disable(logging.NOTSET)
    save("unit_eotf", **d)

    # presets
    names = ["standard_4k", "standard_hdr_pq", "standard_hdr_linear", "standard_fhd", "standard_hmd", "standard_phone"]
    rows = []
    for nme in names:
        ph = fvvdp_display_photometry.load(nme)
        ge = fvvdp_display_geometry.load(nme)
        rows.append([ph.get_peak_luminance(), ph.get_black_level(), ge.get_ppd(), ge.display_size_m[0], ge.display_size_m[1], ge.distance_m])
    save("unit_presets", names=np.array(names), rows=np.array(rows, dtype=np.float64))

    # foveation maps
    ge = fvvdp_display_geometry.load("standard_hmd")
    w, h = 90, 100
    xv = torch.linspace(0.5, w - 0.5, w)
    yv = torch.linspace(0.5, h - 0.5, h)
    xx, yy = torch.meshgrid(xv, yv, indexing="xy")
    vd = ge.pix2view_direction(torch.tensor((w, h)), xx, yy)
    vg = ge.pix2view_direction(torch.tensor((1440, 1600)), torch.as_tensor(300.0 + 0.5), torch.as_tensor(1200.0 + 0.5)).view(2, 1, 1)
    ecc = torch.sqrt(torch.sum((vd - vg) ** 2, dim=0))
    rm = ge.get_resolution_magnification(vd)
    save("unit_foveation", ecc=ecc.numpy(), res_mag=rm.numpy(), band_wh=[w, h], frame_wh=[1440, 1600], gaze=[300.0, 1200.0])


def gen_known_answer():
    """README.md:123-138 / pytorch_examples/ex_simple_image.py: wavy_facade vs Gaussian blur sigma=2."""
    import cv2
    import scipy.ndimage
    I = cv2.imread("/root/reference/example_media/wavy_facade.png", cv2.IMREAD_UNCHANGED)[:, :, ::-1].copy()
    sigma = 2
    # ex_utils.imgaussblur: scipy gaussian_filter applied per colour plane on the uint16 array itself
    Ib = np.stack([scipy.ndimage.gaussian_filter(I[:, :, c], sigma, mode="nearest", truncate=2.0) for c in range(3)], 2)
    out = {}
    for disp in ("standard_fhd", "standard_4k"):
        fv = pyfvvdp.fvvdp(display_name=disp, device=CPU)
        q, st = fv.predict(Ib, I, dim_order="HWC")
        out[disp] = float(q)
        out[disp + "_Q"] = st["Q_per_ch"]
        print(disp, float(q))
    save("known_answer_wavy_facade", **out)


class custom_display_geometry(fvvdp_display_geometry):
    """The custom geometry of pytorch_examples/ex_custom_ppd.py:38-57: ppd falls off with the view angle."""

    def get_ppd(self, view_dir=None):
        if view_dir is None:
            return self.ppd_centre
        view_angle = torch.sqrt(torch.sum((view_dir) ** 2, dim=0, keepdim=False))
        return self.ppd_centre / (view_angle / 20. + 1.)


def gen_widened_cases():
    test, ref = synth_pair_numpy(12, 270, 480)
    # heat-map visualisation: colour maps over the tone-mapped test frame (visualize_diff_map.py)
    for mode in ("threshold", "supra-threshold"):
        fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU, heatmap=mode)
        q, st = fv.predict(torch.tensor(test[:, :, :4]), torch.tensor(ref[:, :, :4]), dim_order="BCFHW", frames_per_second=30)
        hm = st["heatmap"].float().numpy()
        save(f"video_fhd_heatmap_{mode}", jod=float(q), heatmap_sub=hm[0, :, :, ::SY, ::SX], hm_mean=hm.mean(axis=(0, 2, 3, 4)))
    # the same on an image whose luminance range is below the tone-mapping threshold (linear branch of vis_tonemap)
    fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU, heatmap="threshold")
    ti, ri = 0.5 + 0.1 * test[0, :, 0:1], 0.5 + 0.1 * ref[0, :, 0:1]
    q, st = fv.predict(torch.tensor(ti), torch.tensor(ri), dim_order="CFHW")
    hm = st["heatmap"].float().numpy()
    save("image_fhd_heatmap_threshold_lowdr", jod=float(q), heatmap_sub=hm[0, :, :, ::SY, ::SX], hm_mean=hm.mean(axis=(0, 2, 3, 4)))
    # foveated scoring with a custom fvvdp_display_geometry subclass (ex_custom_ppd.py), moving gaze
    gaze = np.stack([np.linspace(0, 479, 6), np.linspace(269, 0, 6)], 1).astype(np.float32)
    geo = custom_display_geometry([480, 270], distance_m=0.6, diagonal_size_inches=24)
    fv = pyfvvdp.fvvdp(display_name="standard_fhd", display_geometry=geo, device=CPU, foveated=True)
    q, st = fv.predict(torch.tensor(test[:, :, :6]), torch.tensor(ref[:, :, :6]), dim_order="BCFHW", frames_per_second=30, fixation_point=gaze)
    save("video_custom_geometry_foveated", jod=float(q), Q_per_ch=st["Q_per_ch"], gaze=gaze, rho_band=st["rho_band"])


def gen_full_size_cases():
    """BASELINE.json frame sizes through the unmodified reference (CPU): configs[1] 1920x1080 on standard_fhd and configs[2]
    3840x2160 on standard_4k, 30 fps, a few frames of the same analytic clip the benchmark uses -- only JOD and Q_per_ch are kept."""
    t, r = synth_pair_numpy(10, 1080, 1920)
    fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU)
    q, st = fv.predict(torch.tensor(t), torch.tensor(r), dim_order="BCFHW", frames_per_second=30)
    save("full_fhd_10f", jod=float(q), Q_per_ch=st["Q_per_ch"], rho_band=st["rho_band"])
    t, r = synth_pair_numpy(9, 2160, 3840)
    fv = pyfvvdp.fvvdp(display_name="standard_4k", device=CPU)
    q, st = fv.predict(torch.tensor(t), torch.tensor(r), dim_order="BCFHW", frames_per_second=30)
    save("full_4k_9f", jod=float(q), Q_per_ch=st["Q_per_ch"], rho_band=st["rho_band"])
    gaze = np.stack([np.linspace(0, 3839, 9), np.linspace(0, 2159, 9)], 1).astype(np.float32)
    fv = pyfvvdp.fvvdp(display_name="standard_hdr_pq", device=CPU, foveated=True)
    q, st = fv.predict(torch.tensor(0.1 + 0.65 * t), torch.tensor(0.1 + 0.65 * r), dim_order="BCFHW", frames_per_second=30, fixation_point=gaze)
    save("full_4k_hdr_pq_foveated_9f", jod=float(q), Q_per_ch=st["Q_per_ch"], gaze=gaze)


def gen_round2_cases():
    """Fixtures added in round 2: the benchmark clip itself (64 frames of 3840x2160, BASELINE configs[2]), the GOG photometry
    (fvvdp_display_model.py:253-279) with a gamma and with its sRGB branch, 120 fps (30-tap temporal window, fvvdp.py:228), and a
    source whose colour space differs from the metric's (the source's RGB->Y weights apply, video_source.py:87,206)."""
    from pyfvvdp.fvvdp_display_model import fvvdp_display_photo_gog
    from pyfvvdp.video_source import fvvdp_video_source_array
    t, r = synth_pair_numpy(8, 270, 480)
    for name, gamma in (("gog_gamma", 2.4), ("gog_srgb", -1)):
        dp = fvvdp_display_photo_gog(300.0, contrast=500, gamma=gamma, E_ambient=100, k_refl=0.005)
        fv = pyfvvdp.fvvdp(display_name="standard_fhd", display_photometry=dp, device=CPU)
        q, st = fv.predict(torch.tensor(t), torch.tensor(r), dim_order="BCFHW", frames_per_second=30)
        save(f"video_{name}", jod=float(q), Q_per_ch=st["Q_per_ch"], gog=[300.0, 500.0, gamma, 100.0, 0.005])
    for (N, H, W) in ((9, 135, 240), (40, 270, 480)):
        t, r = synth_pair_numpy(N, H, W)
        fv = pyfvvdp.fvvdp(display_name="standard_fhd", device=CPU)
        q, st = fv.predict(torch.tensor(t), torch.tensor(r), dim_order="BCFHW", frames_per_second=120)
        save(f"video_120fps_{N}x{H}x{W}", jod=float(q), Q_per_ch=st["Q_per_ch"])
    rng = np.random.default_rng(5)
    t8 = rng.integers(0, 256, (6, 96, 160, 3), dtype=np.uint8)
    r8 = np.clip(t8.astype(np.int32) + rng.integers(-12, 13, t8.shape), 0, 255).astype(np.uint8)
    fv = pyfvvdp.fvvdp(display_name="standard_4k", device=CPU)  # metric left at color_space="sRGB"
    vs = fvvdp_video_source_array(torch.tensor(t8), torch.tensor(r8), 30, dim_order="FHWC", display_photometry=fv.display_photometry,
                                  color_space_name="BT.2020")
    q, st = fv.predict_video_source(vs)
    save("video_source_bt2020_metric_srgb", jod=float(q), Q_per_ch=st["Q_per_ch"], test=t8, ref=r8)
    t, r = synth_pair_numpy(64, 2160, 3840)
    fv = pyfvvdp.fvvdp(display_name="standard_4k", device=CPU)
    q, st = fv.predict(torch.tensor(t), torch.tensor(r), dim_order="BCFHW", frames_per_second=30)
    save("full_4k_64f", jod=float(q), Q_per_ch=st["Q_per_ch"], rho_band=st["rho_band"])


def gen_pu_psnr_cases():
    """PU21-PSNR through the reference's pu_psnr.predict_video_source (pupsnr.py:52-79) with its array video source."""
    from pyfvvdp.video_source import fvvdp_video_source_array
    test, ref = synth_pair_numpy(5, 135, 240)
    out = {}
    m = pyfvvdp.pu_psnr(device=CPU)
    for disp in ("standard_4k", "standard_hdr_pq"):
        t, r = (test, ref) if disp == "standard_4k" else (0.1 + 0.65 * test, 0.1 + 0.65 * ref)
        vs = fvvdp_video_source_array(torch.tensor(t), torch.tensor(r), 30, dim_order="BCFHW", display_photometry=disp)
        q, _ = m.predict_video_source(vs)
        out[disp] = float(q)
    rng = np.random.default_rng(7)
    t8 = rng.integers(0, 256, (3, 40, 56, 3), dtype=np.uint8)
    r8 = np.clip(t8.astype(np.int16) + rng.integers(-6, 7, t8.shape), 0, 255).astype(np.uint8)
    vs = fvvdp_video_source_array(t8, r8, 30, dim_order="FHWC", display_photometry="standard_fhd")
    out["u8_rgb_fhd"] = float(m.predict_video_source(vs)[0])
    save("pu_psnr", test_u8=t8, ref_u8=r8, **out)


def gen_yuv_cases():
    """Raw .yuv clips through the reference's fvvdp_video_source_yuv_file (video_source_yuv.py).  That module imports its
    siblings as top-level modules, so the package directory itself goes on sys.path."""
    import tempfile
    sys.path.insert(0, "/root/reference/pyfvvdp")
    import video_source_yuv as vy
    from fovvideovdp_b200.synthetic import synth_yuv_pair
    # the reference's constructor formats a debug message from two attributes YUVReader never sets
    # (video_source_yuv.py:266, AttributeError); give the class placeholders -- no arithmetic is touched
    vy.YUVReader.color_transfer = "n/a"
    vy.YUVReader.in_pix_fmt = "n/a"
    for (H, W, bits, css, cs, fps, disp) in ((96, 160, 10, "420", "2020", 30, "standard_hdr_pq"), (72, 100, 8, "444", "709", 25, "standard_4k")):
        t, r = synth_yuv_pair(6, H, W, bits, css)
        with tempfile.TemporaryDirectory() as d:
            props = dict(width=W, height=H, bit_depth=bits, color_space=cs, chroma_ss=css, fps=fps)
            ft, fr = os.path.join(d, vy.create_yuv_fname("test", props)), os.path.join(d, vy.create_yuv_fname("ref", props))
            t.tofile(ft)
            r.tofile(fr)
            vs = vy.fvvdp_video_source_yuv_file(ft, fr, display_photometry=disp)
            fv = pyfvvdp.fvvdp(display_name=disp, device=CPU)
            q, st = fv.predict_video_source(vs)
            lum0 = vs.get_test_frame(2, CPU).numpy()[0, 0, 0]
            rgb0 = vs.test_vidr.get_frame_rgb_tensor(2, CPU).numpy()
            save(f"yuv_{bits}b_{css}_{cs}", jod=float(q), Q_per_ch=st["Q_per_ch"], lum_test_f2=lum0, rgb_test_f2=rgb0[::3, ::3],
                 H=H, W=W, bits=bits, fps=fps, frames_per_second=st["frames_per_second"])


def gen_yuv_resize_cases():
    """Full-screen resize of raw .yuv clips (video_source_yuv.py:293-297): every torch.nn.functional.interpolate mode the
    reference's command line offers, up- and down-scaling with non-integer factors; one frame of luminance + the JOD each."""
    import tempfile
    sys.path.insert(0, "/root/reference/pyfvvdp")
    import video_source_yuv as vy
    from fovvideovdp_b200.synthetic import synth_yuv_pair
    vy.YUVReader.color_transfer = "n/a"   # see gen_yuv_cases
    vy.YUVReader.in_pix_fmt = "n/a"
    H, W, bits, css, cs, fps = 96, 160, 10, "420", "2020", 30
    t, r = synth_yuv_pair(4, H, W, bits, css)
    out = dict(H=H, W=W, bits=bits, fps=fps)
    with tempfile.TemporaryDirectory() as d:
        props = dict(width=W, height=H, bit_depth=bits, color_space=cs, chroma_ss=css, fps=fps)
        ft, fr = os.path.join(d, vy.create_yuv_fname("test", props)), os.path.join(d, vy.create_yuv_fname("ref", props))
        t.tofile(ft)
        r.tofile(fr)
        for disp in ("standard_hdr_pq", "standard_4k"):
            for mode in ("nearest", "bilinear", "bicubic", "area"):
                for tag, res in (("up", (200, 130)), ("down", (116, 75))):
                    vs = vy.fvvdp_video_source_yuv_file(ft, fr, display_photometry=disp, full_screen_resize=mode, resize_resolution=res)
                    key = f"{disp}_{mode}_{tag}"
                    out["lum_" + key] = vs.get_test_frame(1, CPU).numpy()[0, 0, 0]
                    if disp == "standard_hdr_pq":
                        q, st = pyfvvdp.fvvdp(display_name=disp, device=CPU).predict_video_source(vs)
                        out["jod_" + key] = float(q)
                        out["Q_" + key] = st["Q_per_ch"].numpy() if hasattr(st["Q_per_ch"], "numpy") else st["Q_per_ch"]
    save("yuv_resize", **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    if len(sys.argv) > 1 and sys.argv[1] == "widen":
        gen_widened_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        gen_full_size_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        gen_round2_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pupsnr":
        gen_pu_psnr_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "yuv":
        gen_yuv_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "yuvresize":
        gen_yuv_resize_cases()
        sys.exit(0)
    gen_unit_cases()
    gen_metric_cases()
    gen_known_answer()
    gen_widened_cases()
    gen_yuv_cases()
    gen_yuv_resize_cases()
    gen_pu_psnr_cases()
    gen_full_size_cases()
    gen_round2_cases()
