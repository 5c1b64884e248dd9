"""Pin oracle/fvvdp_oracle.py against vectors produced by the unmodified reference
(tools/gen_golden.py -> tests/golden/*.npz) and the README known answer."""
import os

import numpy as np
import pytest

from fovvideovdp_b200.synthetic import synth_pair_numpy
from oracle import fvvdp_oracle as O

SY, SX = 5, 7
JOD_RTOL = 1e-5          # oracle (numpy fp32) vs reference (torch fp32): pure rounding-order noise


def sub(a):
    return a if a.size <= 40000 else a[..., ::SY, ::SX]


def test_pyramid_reduce_expand(golden):
    g = golden("unit_pyramid")
    for i in range(int(g["n"])):
        x = g[f"x{i}"]
        red = O.gausspyr_reduce(x)
        np.testing.assert_allclose(red, g[f"red{i}"], rtol=0, atol=2e-5 * 100)
        ex = O.gausspyr_expand(g[f"red{i}"], x.shape[-2:])
        np.testing.assert_allclose(ex, g[f"exp{i}"], rtol=0, atol=2e-5 * 100)


def test_parity_quirk_is_reproduced(golden):
    """fvvdp_lpyr_dec.py:202 keys the column edge term on the ROW count: (135,240) and (136,241) differ
    from a clean symmetric-pad reduce in the last column only; the oracle must follow the reference."""
    g = golden("unit_pyramid")
    x = g["x0"]  # (135, 240)
    clean = O._reduce_axis(O._reduce_axis(x, -2, True), -1, False)
    assert np.abs(clean[..., :-1] - g["red0"][..., :-1]).max() < 1e-3
    assert np.abs(clean[..., -1] - g["red0"][..., -1]).max() > 1.0


def test_contrast_pyramid(golden):
    g = golden("unit_contrast_pyr")
    h = int(g["height"])
    bands, lbkg, _ = O.contrast_pyramid(g["x"][:, 0], h)
    for i in range(h + 1):
        np.testing.assert_allclose(bands[i], g[f"band{i}"][:, 0], rtol=2e-5, atol=2e-4)
    for i in range(h):
        np.testing.assert_allclose(lbkg[i], g[f"lbkg{i}"][0, 0], rtol=1e-5, atol=1e-5)


def test_pyramid_layout(golden):
    for row in golden("unit_pyr_layout")["rows"]:
        W, H, ppd, height = int(row[0]), int(row[1]), row[2], int(row[3])
        h, f = O.pyramid_layout(W, H, ppd)
        assert h == height
        np.testing.assert_allclose(f, row[4:4 + height + 1], rtol=1e-12)


def test_csf_and_masking(golden):
    g = golden("unit_csf_masking")
    S0 = O.csf_sensitivity(g["rho"], 0, g["L"], g["ecc"])
    S5 = O.csf_sensitivity(g["rho"], 1, g["L"], g["ecc"])
    np.testing.assert_allclose(S0, g["S0"], rtol=2e-5)
    np.testing.assert_allclose(S5, g["S5"], rtol=2e-5)
    p = O.metric_data()["parameters"]
    for cc, key in ((0, "D0"), (1, "D1")):
        D = O.masking(g["T"], g["R"], g["Smul"], cc, p)
        np.testing.assert_allclose(D, g[key], rtol=2e-4, atol=1e-12)


def test_temporal_filters(golden):
    g = golden("unit_temporal_filters")
    for fps in (24, 25, 30, 50, 60, 120, 12.5):
        F = O.temporal_filters(fps, O.filter_len(fps))
        np.testing.assert_allclose(F, g[f"F_{fps}"], rtol=1e-4, atol=2e-7)


def test_eotf(golden):
    g = golden("unit_eotf")
    V = g["V"]
    for kind in ("sRGB", "gamma", "PQ", "linear"):
        Yp = 1500 if kind in ("PQ", "linear") else 200
        Yb = 250 / np.pi * 0.005 + Yp / 1000
        assert abs(Yb - float(g[kind + "_black"])) < 1e-9
        Vin = V * 2000 if kind == "linear" else V
        L = O.eotf_forward(Vin, dict(kind=kind, Y_peak=Yp, Y_black=Yb, gamma=2.2))
        np.testing.assert_allclose(L, g[kind], rtol=2e-5, atol=1e-6)
    L = O.eotf_forward(V * 2000, dict(kind="absolute", L_min=0.01, L_max=1000))
    np.testing.assert_allclose(L, g["absolute"], rtol=1e-7)


def test_presets(golden):
    g = golden("unit_presets")
    for name, row in zip(g["names"], g["rows"]):
        ph, ge = O.photometry_from_preset(str(name)), O.geometry_from_preset(str(name))
        got = [ph["Y_peak"], ph["Y_black"], ge["ppd_centre"], ge["display_size_m"][0], ge["display_size_m"][1], ge["distance_m"]]
        np.testing.assert_allclose(got, row, rtol=1e-12)


def test_foveation_maps(golden):
    g = golden("unit_foveation")
    geo = O.geometry_from_preset("standard_hmd")
    w, h = [int(v) for v in g["band_wh"]]
    fw, fh = [int(v) for v in g["frame_wh"]]
    ecc, rm = O.foveation_maps(geo, (h, w), (fh, fw), g["gaze"])
    np.testing.assert_allclose(ecc, g["ecc"], rtol=1e-4, atol=1e-4)
    # the reference forms tan(a+delta)-tan(a) in fp32: cancellation noise ~1e-3 relative is inherent
    np.testing.assert_allclose(rm, g["res_mag"], rtol=5e-3)


def _check_taps(g, taps, n_bands, temp_ch, band_atol=1e-3):
    np.testing.assert_allclose(sub(taps["R"]), g["R"], rtol=1e-5, atol=1e-4)
    for cc in range(temp_ch):
        for bb in range(n_bands):
            # north-star tolerance on intermediate band contrasts: <= 1e-3 max-abs
            np.testing.assert_allclose(sub(taps["T_f"][(bb, cc)]), g[f"T_f_{bb}_{cc}"], rtol=0, atol=band_atol)
            np.testing.assert_allclose(sub(taps["R_f"][(bb, cc)]), g[f"R_f_{bb}_{cc}"], rtol=0, atol=band_atol)
            np.testing.assert_allclose(sub(taps["S"][(bb, cc)]), g[f"S_{bb}_{cc}"], rtol=1e-4)
            gD = g[f"D_{bb}_{cc}"]
            np.testing.assert_allclose(sub(taps["D"][(bb, cc)]), gD, rtol=5e-3, atol=1e-5 + 1e-4 * float(gD.max()))
        np.testing.assert_allclose(sub(taps["L_bkg"][bb]), g[f"L_bkg_{bb}"], rtol=1e-5)


def _check_q(Q, gQ, tol=1e-4):
    # per-band pooled values, relative to the largest value of that temporal channel (bands whose
    # T-R difference is pure cancellation noise are not meaningful on their own scale)
    scale = np.maximum(np.abs(gQ).max(axis=(0, 2), keepdims=True), 1e-6)
    assert (np.abs(Q - gQ) / scale).max() < tol


def test_video_fhd_replicate(golden):
    g = golden("video_fhd_replicate")
    t, r = synth_pair_numpy(12, 270, 480)
    jod, st = O.predict(t, r, frames_per_second=30, display_name="standard_fhd", tap_frame=int(g["tap_frame"]))
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    assert abs(float(g["jod"]) - 9.447366) < 2e-6
    _check_q(st["Q_per_ch"], g["Q_per_ch"])
    np.testing.assert_allclose(st["rho_band"], g["rho_band"], rtol=1e-12)
    _check_taps(g, st["taps"], 6, 2)


@pytest.mark.parametrize("pad", ["pingpong", "circular"])
def test_video_padding_modes(golden, pad):
    g = golden(f"video_fhd_{pad}")
    t, r = synth_pair_numpy(12, 270, 480)
    jod, st = O.predict(t, r, frames_per_second=30, display_name="standard_fhd", temp_padding=pad)
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    _check_q(st["Q_per_ch"], g["Q_per_ch"])


@pytest.mark.parametrize("fps", [25, 60])
@pytest.mark.parametrize("pad", ["replicate", "pingpong", "circular"])
def test_short_clip_padding(golden, fps, pad):
    g = golden(f"video_short_{fps}fps_{pad}")
    t, r = synth_pair_numpy(5, 270, 480)
    jod, st = O.predict(t, r, frames_per_second=fps, display_name="standard_fhd", temp_padding=pad)
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    _check_q(st["Q_per_ch"], g["Q_per_ch"])


def test_image(golden):
    g = golden("image_fhd")
    t, r = synth_pair_numpy(1, 270, 480)
    jod, st = O.predict(t[0, :, 0:1], r[0, :, 0:1], dim_order="CFHW", display_name="standard_fhd", tap_frame=0)
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    assert abs(float(g["jod"]) - 9.876795) < 2e-6  # App. D, K2
    assert np.all(st["Q_per_ch"][:, 1] == 0)
    _check_taps(g, st["taps"], 6, 1)


def test_foveated_hdr_pq(golden):
    g = golden("video_hdrpq_foveated")
    t, r = synth_pair_numpy(12, 270, 480)
    jod, st = O.predict(0.1 + 0.65 * t, 0.1 + 0.65 * r, frames_per_second=30, display_name="standard_hdr_pq",
                        foveated=True, fixation_point=g["gaze"], tap_frame=int(g["tap_frame"]))
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < 2e-5
    # the reference forms tan(a+delta)-tan(a) in fp32 (cancellation noise ~1e-3 in rho): looser per-band bound
    _check_q(st["Q_per_ch"], g["Q_per_ch"], tol=1e-3)


def test_foveated_hmd_fixed_gaze(golden):
    g = golden("video_hmd_foveated_fixed")
    t, r = synth_pair_numpy(4, 270, 480)
    jod, st = O.predict(t, r, frames_per_second=30, display_name="standard_hmd", foveated=True,
                        fixation_point=np.array([100.0, 50.0], np.float32))
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < 2e-5
    _check_q(st["Q_per_ch"], g["Q_per_ch"], tol=1e-3)


def test_heatmap_raw(golden):
    g = golden("video_fhd_heatmap_raw")
    t, r = synth_pair_numpy(4, 270, 480)
    jod, st = O.predict(t, r, frames_per_second=30, display_name="standard_fhd", heatmap="raw")
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    hm = st["heatmap"].astype(np.float32)
    np.testing.assert_allclose(hm[0, 0, :, ::SY, ::SX], g["heatmap_sub"], rtol=2e-3, atol=2e-3)  # fp16 storage
    assert abs(hm.mean() - float(g["hm_mean"])) < 1e-4


@pytest.mark.parametrize("mode", ["threshold", "supra-threshold"])
def test_heatmap_colour_maps(golden, mode):
    """visualize_diff_map (visualize_diff_map.py:58-107): colour map of the difference map over the tone-mapped
    sustained test frame."""
    g = golden(f"video_fhd_heatmap_{mode}")
    t, r = synth_pair_numpy(4, 270, 480)
    jod, st = O.predict(t, r, frames_per_second=30, display_name="standard_fhd", heatmap=mode)
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    hm = st["heatmap"].astype(np.float32)
    assert hm.shape == (1, 3, 4, 270, 480)
    np.testing.assert_allclose(hm[0, :, :, ::SY, ::SX], g["heatmap_sub"], rtol=2e-3, atol=2e-3)  # fp16 storage
    np.testing.assert_allclose(hm.mean(axis=(0, 2, 3, 4)), g["hm_mean"], atol=2e-4)


def test_heatmap_colour_map_low_dynamic_range(golden):
    """Luminance range of the context image below 0.6 log units: the linear branch of vis_tonemap (:32-34)."""
    g = golden("image_fhd_heatmap_threshold_lowdr")
    t, r = synth_pair_numpy(1, 270, 480)
    ti, ri = 0.5 + 0.1 * t[0, :, 0:1], 0.5 + 0.1 * r[0, :, 0:1]
    jod, st = O.predict(ti, ri, dim_order="CFHW", display_name="standard_fhd", heatmap="threshold")
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    hm = st["heatmap"].astype(np.float32)
    np.testing.assert_allclose(hm[0, :, :, ::SY, ::SX], g["heatmap_sub"], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(hm.mean(axis=(0, 2, 3, 4)), g["hm_mean"], atol=2e-4)


def custom_geometry_oracle():
    """numpy twin of the fvvdp_display_geometry subclass in pytorch_examples/ex_custom_ppd.py:38-57
    (ppd = ppd_centre / (view_angle / 20 + 1)) on a 480x270, 24-inch display seen from 0.6 m."""
    geo = O.geometry((480, 270), distance_m=0.6, diagonal_size_inches=24)

    def res_mag(geo_, vx, vy):
        va = np.sqrt(vx * vx + vy * vy).astype(np.float32)
        return (np.float32(geo_["ppd_centre"]) / (va / np.float32(20.0) + np.float32(1.0)) / np.float32(geo_["ppd_centre"])).astype(np.float32)

    geo["resolution_magnification"] = res_mag
    return geo


def test_custom_geometry_foveated(golden):
    g = golden("video_custom_geometry_foveated")
    t, r = synth_pair_numpy(12, 270, 480)
    jod, st = O.predict(t[:, :, :6], r[:, :, :6], frames_per_second=30, display_name="standard_fhd", geometry_=custom_geometry_oracle(),
                        foveated=True, fixation_point=g["gaze"])
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    np.testing.assert_allclose(st["rho_band"], g["rho_band"], rtol=1e-6)
    _check_q(st["Q_per_ch"], g["Q_per_ch"], tol=1e-3)


def test_full_hd_against_reference(golden):
    """BASELINE configs[1] frame size: 1920x1080, standard_fhd, 30 fps, the first 10 frames of the benchmark's analytic clip."""
    g = golden("full_fhd_10f")
    t, r = synth_pair_numpy(10, 1080, 1920)
    jod, st = O.predict(t, r, frames_per_second=30, display_name="standard_fhd")
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    _check_q(st["Q_per_ch"], g["Q_per_ch"])


def test_pu_psnr(golden):
    """PU21-PSNR (pupsnr.py:52-79, utils.py:157-202) against the reference on sRGB, PQ and uint8 RGB content."""
    g = golden("pu_psnr")
    t, r = synth_pair_numpy(5, 135, 240)
    assert abs(O.pu_psnr(t, r, display_name="standard_4k") - float(g["standard_4k"])) < 2e-3
    assert abs(O.pu_psnr(0.1 + 0.65 * t, 0.1 + 0.65 * r, display_name="standard_hdr_pq") - float(g["standard_hdr_pq"])) < 2e-3
    assert abs(O.pu_psnr(g["test_u8"], g["ref_u8"], dim_order="FHWC", display_name="standard_fhd") - float(g["u8_rgb_fhd"])) < 2e-3


YUV_CASES = [("yuv_10b_420_2020", "420", "2020", "standard_hdr_pq"), ("yuv_8b_444_709", "444", "709", "standard_4k")]


@pytest.mark.parametrize("case", YUV_CASES)
def test_yuv_video_source(golden, case):
    """Raw planar .yuv clips (video_source_yuv.py:157-302): fixed2float, bilinear chroma upsampling, ycbcr2rgb, display
    model, metric."""
    from fovvideovdp_b200.synthetic import synth_yuv_pair
    name, css, cs, disp = case
    g = golden(name)
    H, W, bits = int(g["H"]), int(g["W"]), int(g["bits"])
    ny, nc = H * W, (H * W // 4 if css == "420" else H * W)
    ch, cw = (H // 2, W // 2) if css == "420" else (H, W)

    def rgb_clip(frames):
        return np.stack([O.yuv_frame_rgb(f[:ny].reshape(H, W), f[ny:ny + nc].reshape(ch, cw), f[ny + nc:].reshape(ch, cw), bits, css, cs)
                         for f in frames], 0)

    t, r = synth_yuv_pair(6, H, W, bits, css)
    rt, rr = rgb_clip(t), rgb_clip(r)
    np.testing.assert_allclose(rt[2][::3, ::3], g["rgb_test_f2"], atol=2e-6)
    photo = O.photometry_from_preset(disp)
    cspace = "BT.2020" if cs == "2020" else "sRGB"
    lum = O.frame_luminance(np.transpose(rt[2], (2, 0, 1)), photo, O.metric_data()["rgb2y"][cspace])
    np.testing.assert_allclose(lum, g["lum_test_f2"], rtol=3e-4, atol=1e-4)  # PQ amplifies the 1e-6 differences of the RGB stage
    jod, st = O.predict(rt, rr, dim_order="FHWC", frames_per_second=float(g["fps"]), display_name=disp, color_space=cspace)
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    _check_q(st["Q_per_ch"], g["Q_per_ch"])


def test_yuv_full_screen_resize(golden):
    """--full-screen-resize of raw .yuv clips (video_source_yuv.py:293-297): the oracle's restatement of the four interpolate
    modes against the reference's frames, up- and down-scaling by non-integer factors, and one score."""
    from fovvideovdp_b200.synthetic import synth_yuv_pair
    g = golden("yuv_resize")
    H, W, bits, fps = int(g["H"]), int(g["W"]), int(g["bits"]), float(g["fps"])
    ny, nc = H * W, H * W // 4
    t, r = synth_yuv_pair(4, H, W, bits, "420")
    rgb = lambda f: O.yuv_frame_rgb(f[:ny].reshape(H, W), f[ny:ny + nc].reshape(H // 2, W // 2), f[ny + nc:].reshape(H // 2, W // 2), bits, "420", "2020")
    frame = rgb(t[1])
    for disp in ("standard_hdr_pq", "standard_4k"):
        photo = O.photometry_from_preset(disp)
        for mode in ("nearest", "bilinear", "bicubic", "area"):
            for tag, res in (("up", (200, 130)), ("down", (116, 75))):
                out = O.resize_rgb(frame, res[0], res[1], mode)
                assert out.shape == (res[1], res[0], 3)
                lum = O.frame_luminance(np.transpose(out, (2, 0, 1)), photo, O.metric_data()["rgb2y"]["BT.2020"])
                np.testing.assert_allclose(lum, g[f"lum_{disp}_{mode}_{tag}"], rtol=3e-4, atol=1e-4, err_msg=f"{disp} {mode} {tag}")
    rt = np.stack([O.resize_rgb(rgb(f), 200, 130, "bicubic") for f in t], 0)
    rr = np.stack([O.resize_rgb(rgb(f), 200, 130, "bicubic") for f in r], 0)
    jod, st = O.predict(rt, rr, dim_order="FHWC", frames_per_second=fps, display_name="standard_hdr_pq", color_space="BT.2020")
    assert abs(jod - float(g["jod_standard_hdr_pq_bicubic_up"])) / float(g["jod_standard_hdr_pq_bicubic_up"]) < JOD_RTOL
    _check_q(st["Q_per_ch"], g["Q_standard_hdr_pq_bicubic_up"])


@pytest.mark.parametrize("hw", [(135, 240), (136, 241), (67, 97), (64, 64)])
def test_odd_sizes(golden, hw):
    H, W = hw
    g = golden(f"video_4k_{H}x{W}")
    t, r = synth_pair_numpy(3, H, W)
    jod, st = O.predict(t, r, frames_per_second=24, display_name="standard_4k", tap_frame=int(g["tap_frame"]))
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    _check_q(st["Q_per_ch"], g["Q_per_ch"])
    _check_taps(g, st["taps"], st["Q_per_ch"].shape[0], 2)


def test_u8_rgb_fhwc(golden):
    g = golden("video_u8_rgb_fhwc")
    jod, st = O.predict(g["test"], g["ref"], dim_order="FHWC", frames_per_second=30, display_name="standard_fhd")
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    _check_q(st["Q_per_ch"], g["Q_per_ch"])


def test_u16_rgb_gamma_bt2020(golden):
    g = golden("image_u16_rgb_gamma_bt2020")
    Yb = 100 / np.pi * 0.005 + 400 / 2000
    jod, st = O.predict(g["test"], g["ref"], dim_order="HWC", display_name="standard_4k", color_space="BT.2020",
                        photometry=dict(kind="gamma", Y_peak=400.0, Y_black=Yb, gamma=2.4))
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    _check_q(st["Q_per_ch"], g["Q_per_ch"])


def test_absolute_and_linear(golden):
    t2, r2 = synth_pair_numpy(3, 64, 64)
    ta, ra = (t2 * 300 + 0.001).astype(np.float32), (r2 * 300 + 0.001).astype(np.float32)
    g = golden("video_absolute")
    jod, st = O.predict(ta, ra, frames_per_second=30, display_name="standard_4k", photometry=dict(kind="absolute", L_min=0.01, L_max=4000))
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    g = golden("video_hdr_linear")
    jod, st = O.predict(ta, ra, frames_per_second=30, display_name="standard_hdr_linear")
    assert abs(jod - float(g["jod"])) / float(g["jod"]) < JOD_RTOL
    _check_q(st["Q_per_ch"], g["Q_per_ch"])


WAVY = "/root/reference/example_media/wavy_facade.png"


@pytest.mark.skipif(not os.path.isfile(WAVY), reason="reference media only exists in the build container")
def test_readme_known_answer(golden):
    """README.md:123-138: wavy_facade vs Gaussian blur sigma=2 on standard_4k -> 8.693 JOD."""
    import cv2
    import scipy.ndimage
    g = golden("known_answer_wavy_facade")
    I = cv2.imread(WAVY, cv2.IMREAD_UNCHANGED)[:, :, ::-1].copy()
    Ib = np.stack([scipy.ndimage.gaussian_filter(I[:, :, c], 2, mode="nearest", truncate=2.0) for c in range(3)], 2)
    jod, st = O.predict(Ib, I, dim_order="HWC", display_name="standard_4k")
    assert abs(jod - 8.693) < 5e-4
    assert abs(jod - float(g["standard_4k"])) / jod < JOD_RTOL
    _check_q(st["Q_per_ch"], g["standard_4k_Q"])
    jod, st = O.predict(Ib, I, dim_order="HWC", display_name="standard_fhd")
    assert abs(jod - float(g["standard_fhd"])) / jod < JOD_RTOL
