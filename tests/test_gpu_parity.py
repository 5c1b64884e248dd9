"""Parity of the CUDA path (through the C ABI, via fovvideovdp_b200.fvvdp) against
  (a) the golden vectors produced by the unmodified reference (tests/golden/*.npz, tools/gen_golden.py),
  (b) the CPU oracle (oracle/fvvdp_oracle.py) on the same seeded inputs,
  (c) size-independent properties at the benchmark's full frame sizes.

Tolerances (BASELINE.json north_star): <= 1e-4 relative on the final JOD, <= 1e-3 max-abs on intermediate band
contrasts (fp32).
"""
import numpy as np
import pytest
import torch

from fovvideovdp_b200.synthetic import synth_pair_numpy, synth_pair_torch

pytestmark = pytest.mark.gpu

JOD_RTOL = 1e-4
BAND_ATOL = 1e-3
SY, SX = 5, 7


@pytest.fixture(scope="module")
def fv_mod():
    import fovvideovdp_b200 as m
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device"
    return m


@pytest.fixture(scope="module")
def oracle():
    from oracle import fvvdp_oracle
    return fvvdp_oracle


def sub(a):
    return a if a.size <= 40000 else a[..., ::SY, ::SX]


def check_jod(jod, want, rtol=JOD_RTOL):
    jod = float(jod)
    assert abs(jod - float(want)) / abs(float(want)) < rtol, (jod, float(want))


def check_q(Q, gQ, tol=2e-4):
    scale = np.maximum(np.abs(gQ).max(axis=(0, 2), keepdims=True), 1e-6)
    err = (np.abs(Q - gQ) / scale).max()
    assert err < tol, err


def check_taps(fv, g, n_bands, temp_ch, frame_in_block, N=None):
    from fovvideovdp_b200 import _native as nt
    R = fv.read_tap(nt.TAP_R, 0, frame_in_block).cpu().numpy()
    np.testing.assert_allclose(sub(R), g["R"], rtol=2e-5, atol=2e-4)
    for bb in range(n_bands):
        Cb = fv.read_tap(nt.TAP_CONTRAST, bb, frame_in_block).cpu().numpy()
        Sb = fv.read_tap(nt.TAP_S, bb, frame_in_block).cpu().numpy()
        Db = fv.read_tap(nt.TAP_D, bb, frame_in_block).cpu().numpy()
        Lb = fv.read_tap(nt.TAP_LBKG, bb, frame_in_block).cpu().numpy()[0]
        np.testing.assert_allclose(sub(Lb), g[f"L_bkg_{bb}"], rtol=2e-5)
        for cc in range(temp_ch):
            np.testing.assert_allclose(sub(Cb[2 * cc + 0]), g[f"T_f_{bb}_{cc}"], rtol=0, atol=BAND_ATOL)
            np.testing.assert_allclose(sub(Cb[2 * cc + 1]), g[f"R_f_{bb}_{cc}"], rtol=0, atol=BAND_ATOL)
            np.testing.assert_allclose(sub(Sb[cc]), g[f"S_{bb}_{cc}"], rtol=2e-4)
            gD = g[f"D_{bb}_{cc}"]
            np.testing.assert_allclose(sub(Db[cc]), gD, rtol=5e-3, atol=1e-5 + 2e-4 * float(gD.max()))


# ---------------------------------------------------------------------------------------------- golden vectors
def test_video_fhd_replicate_with_taps(fv_mod, golden):
    g = golden("video_fhd_replicate")
    t, r = synth_pair_numpy(12, 270, 480)
    fv = fv_mod.fvvdp(display_name="standard_fhd")
    fv.debug_taps = True
    jod, st = fv.predict(t, r, frames_per_second=30)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])
    np.testing.assert_allclose(st["rho_band"], g["rho_band"], rtol=1e-12)
    assert st["N_frames"] == 12 and st["width"] == 480 and st["height"] == 270 and st["frames_per_second"] == 30
    check_taps(fv, g, 6, 2, int(g["tap_frame"]))


@pytest.mark.parametrize("pad", ["pingpong", "circular"])
def test_video_padding_modes(fv_mod, golden, pad):
    g = golden(f"video_fhd_{pad}")
    t, r = synth_pair_numpy(12, 270, 480)
    jod, st = fv_mod.fvvdp(display_name="standard_fhd", temp_padding=pad).predict(t, r, frames_per_second=30)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])


@pytest.mark.parametrize("fps", [25, 60])
@pytest.mark.parametrize("pad", ["replicate", "pingpong", "circular"])
def test_short_clip_padding(fv_mod, golden, fps, pad):
    """5-frame clips: the temporal window (7 / 15 taps) is longer than the clip."""
    g = golden(f"video_short_{fps}fps_{pad}")
    t, r = synth_pair_numpy(5, 270, 480)
    jod, st = fv_mod.fvvdp(display_name="standard_fhd", temp_padding=pad).predict(t, r, frames_per_second=fps)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])


@pytest.mark.parametrize("fps", [50, 120])
def test_high_frame_rates_vs_oracle(fv_mod, oracle, fps):
    """50 fps: 13 taps -> the 16-frame on-chip ring (512-thread kernels); 120 fps: 30 taps -> the general kernels."""
    t, r = synth_pair_numpy(9, 135, 240)
    jod, st = fv_mod.fvvdp(display_name="standard_fhd").predict(t, r, frames_per_second=fps)
    want, wst = oracle.predict(t, r, frames_per_second=fps, display_name="standard_fhd")
    check_jod(jod, want)
    check_q(st["Q_per_ch"], wst["Q_per_ch"])


def test_image_with_taps(fv_mod, golden):
    g = golden("image_fhd")
    t, r = synth_pair_numpy(1, 270, 480)
    fv = fv_mod.fvvdp(display_name="standard_fhd")
    fv.debug_taps = True
    jod, st = fv.predict(t[0, :, 0:1], r[0, :, 0:1], dim_order="CFHW")
    check_jod(jod, g["jod"])
    assert st["Q_per_ch"].shape == (6, 2, 1) and np.all(st["Q_per_ch"][:, 1] == 0)
    check_taps(fv, g, 6, 1, 0)


def test_foveated_hdr_pq(fv_mod, golden):
    g = golden("video_hdrpq_foveated")
    t, r = synth_pair_numpy(12, 270, 480)
    fv = fv_mod.fvvdp(display_name="standard_hdr_pq", foveated=True)
    jod, st = fv.predict(0.1 + 0.65 * t, 0.1 + 0.65 * r, frames_per_second=30, fixation_point=g["gaze"])
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"], tol=1e-3)  # the reference's own res_mag carries ~1e-3 fp32 cancellation noise


def test_foveated_hmd_fixed_gaze(fv_mod, golden):
    g = golden("video_hmd_foveated_fixed")
    t, r = synth_pair_numpy(4, 270, 480)
    fv = fv_mod.fvvdp(display_name="standard_hmd", foveated=True)
    jod, st = fv.predict(t, r, frames_per_second=30, fixation_point=torch.tensor([100.0, 50.0]))
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"], tol=1e-3)


def test_heatmap_raw(fv_mod, golden):
    g = golden("video_fhd_heatmap_raw")
    t, r = synth_pair_numpy(4, 270, 480)
    jod, st = fv_mod.fvvdp(display_name="standard_fhd", heatmap="raw").predict(t, r, frames_per_second=30)
    check_jod(jod, g["jod"])
    hm = st["heatmap"]
    assert hm.dtype == torch.float16 and tuple(hm.shape) == (1, 1, 4, 270, 480) and hm.device.type == "cpu"
    hm = hm.float().numpy()
    np.testing.assert_allclose(hm[0, 0, :, ::SY, ::SX], g["heatmap_sub"], rtol=3e-3, atol=3e-3)
    assert abs(hm.mean() - float(g["hm_mean"])) < 2e-4


@pytest.mark.parametrize("mode", ["threshold", "supra-threshold"])
def test_heatmap_colour_maps(fv_mod, golden, mode):
    """heatmap="threshold" / "supra-threshold": colour map over the tone-mapped sustained test frame
    (visualize_diff_map.py:58-107) -> (1,3,N,H,W) fp16 on the CPU.  fp16 storage + histogram-bin jitter of log values."""
    g = golden(f"video_fhd_heatmap_{mode}")
    t, r = synth_pair_numpy(4, 270, 480)
    jod, st = fv_mod.fvvdp(display_name="standard_fhd", heatmap=mode).predict(t, r, frames_per_second=30)
    check_jod(jod, g["jod"])
    hm = st["heatmap"]
    assert hm.dtype == torch.float16 and tuple(hm.shape) == (1, 3, 4, 270, 480) and hm.device.type == "cpu"
    hm = hm.float().numpy()
    np.testing.assert_allclose(hm[0, :, :, ::SY, ::SX], g["heatmap_sub"], rtol=3e-3, atol=3e-3)
    np.testing.assert_allclose(hm.mean(axis=(0, 2, 3, 4)), g["hm_mean"], atol=3e-4)


@pytest.mark.parametrize("mode", ["raw", "threshold"])
def test_heatmap_does_not_depend_on_the_block_cut(fv_mod, mode):
    """Heat maps of a clip scored in blocks of 3 frames (difference-map pyramids and context frames are per block) equal
    those of the clip scored in one block; the fixation point may be a float64 torch tensor."""
    t, r = synth_pair_numpy(8, 135, 240)
    gaze = torch.tensor(np.stack([np.linspace(10, 200, 8), np.linspace(5, 120, 8)], 1), dtype=torch.float64)
    one, s1 = fv_mod.fvvdp(display_name="standard_fhd", heatmap=mode, foveated=True, block_frames=8).predict(t, r, frames_per_second=30, fixation_point=gaze)
    cut, s3 = fv_mod.fvvdp(display_name="standard_fhd", heatmap=mode, foveated=True, block_frames=3).predict(t, r, frames_per_second=30, fixation_point=gaze)
    assert float(one) == float(cut) and np.array_equal(s1["Q_per_ch"], s3["Q_per_ch"])
    assert s1["heatmap"].shape == (1, 1 if mode == "raw" else 3, 8, 135, 240)
    assert torch.equal(s1["heatmap"], s3["heatmap"])


def test_heatmap_colour_map_low_dynamic_range(fv_mod, golden):
    g = golden("image_fhd_heatmap_threshold_lowdr")
    t, r = synth_pair_numpy(1, 270, 480)
    ti, ri = 0.5 + 0.1 * t[0, :, 0:1], 0.5 + 0.1 * r[0, :, 0:1]
    jod, st = fv_mod.fvvdp(display_name="standard_fhd", heatmap="threshold").predict(ti, ri, dim_order="CFHW")
    check_jod(jod, g["jod"])
    hm = st["heatmap"].float().numpy()
    assert hm.shape == (1, 3, 1, 270, 480)
    np.testing.assert_allclose(hm[0, :, :, ::SY, ::SX], g["heatmap_sub"], rtol=3e-3, atol=3e-3)
    np.testing.assert_allclose(hm.mean(axis=(0, 2, 3, 4)), g["hm_mean"], atol=3e-4)


def test_heatmap_colour_map_60fps_general_path(fv_mod, oracle):
    """15-tap window (general kernels): the context frame comes from the materialised temporal channels."""
    t, r = synth_pair_numpy(3, 135, 240)
    jod, st = fv_mod.fvvdp(display_name="standard_fhd", heatmap="supra-threshold").predict(t, r, frames_per_second=60)
    want, wst = oracle.predict(t, r, frames_per_second=60, display_name="standard_fhd", heatmap="supra-threshold")
    check_jod(jod, want)
    np.testing.assert_allclose(st["heatmap"].float().numpy(), wst["heatmap"].astype(np.float32), rtol=3e-3, atol=3e-3)


@pytest.mark.parametrize("case", [("yuv_10b_420_2020", "420", "2020", "standard_hdr_pq"), ("yuv_8b_444_709", "444", "709", "standard_4k")])
def test_yuv_video_source(fv_mod, golden, tmp_path, case):
    """fvvdp_video_source_yuv_file on raw .yuv files: planes uploaded as stored, one conversion kernel per frame
    (fvvdp_b200_yuv_to_luminance), metric through predict_video_source()."""
    from fovvideovdp_b200 import video_source_yuv as vy
    from fovvideovdp_b200.synthetic import synth_yuv_pair
    name, css, cs, disp = case
    g = golden(name)
    H, W, bits = int(g["H"]), int(g["W"]), int(g["bits"])
    t, r = synth_yuv_pair(6, H, W, bits, css)
    props = dict(width=W, height=H, bit_depth=bits, color_space=cs, chroma_ss=css, fps=float(g["fps"]))
    ft, fr = str(tmp_path / vy.create_yuv_fname("test", props)), str(tmp_path / vy.create_yuv_fname("ref", props))
    t.tofile(ft)
    r.tofile(fr)
    assert vy.decode_video_props(ft) == dict(props, fps=float(g["fps"]))
    vs = vy.fvvdp_video_source_yuv_file(ft, fr, display_photometry=disp)
    assert list(vs.get_video_size()) == [H, W, 6] and vs.get_frames_per_second() == float(g["fps"])
    dev = torch.device("cuda:0")
    rgb = vs.test_vidr.get_frame_rgb_tensor(2, dev).cpu().numpy()
    np.testing.assert_allclose(rgb[::3, ::3], g["rgb_test_f2"], atol=2e-6)
    lum = vs.get_test_frame(2, dev)
    assert tuple(lum.shape) == (1, 1, 1, H, W)
    np.testing.assert_allclose(lum.cpu().numpy()[0, 0, 0], g["lum_test_f2"], rtol=5e-4, atol=2e-4)  # PQ amplifies the 1e-6 differences of the RGB stage
    fv = fv_mod.fvvdp(display_name=disp)
    jod, st = fv.predict_video_source(vs)   # block path: raw frames to the device, one conversion launch per block (score_block_yuv)
    assert fv.last_run["h2d_bytes"] == 2 * 6 * t.shape[1] * t.itemsize
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])
    # resized clip (resize kernel; parity: test_yuv_full_screen_resize)
    vs2 = vy.fvvdp_video_source_yuv_file(ft, fr, display_photometry=disp, full_screen_resize="bilinear", resize_resolution=(W * 2, H * 2))
    jod2, st2 = fv.predict_video_source(vs2)
    assert st2["width"] == 2 * W and st2["height"] == 2 * H and 0 < float(jod2) <= 10


def test_yuv_full_screen_resize(fv_mod, golden, tmp_path):
    """--full-screen-resize of raw .yuv clips (video_source_yuv.py:293-297): the four interpolate modes, up- and down-scaling by
    non-integer factors, done inside the conversion kernels (per frame: fvvdp_b200_yuv_to_luminance, per block:
    fvvdp_b200_score_block_yuv).  Checked against the reference's frames and scores (tests/golden/yuv_resize.npz) and, for the
    resampled R'G'B' itself, against torch.nn.functional.interpolate on the kernel's own unresized R'G'B' (fp32 reference of
    the same op)."""
    from fovvideovdp_b200 import video_source_yuv as vy
    from fovvideovdp_b200.synthetic import synth_yuv_pair
    g = golden("yuv_resize")
    H, W, bits, fps = int(g["H"]), int(g["W"]), int(g["bits"]), float(g["fps"])
    t, r = synth_yuv_pair(4, H, W, bits, "420")
    props = dict(width=W, height=H, bit_depth=bits, color_space="2020", chroma_ss="420", fps=fps)
    ft, fr = str(tmp_path / vy.create_yuv_fname("test", props)), str(tmp_path / vy.create_yuv_fname("ref", props))
    t.tofile(ft)
    r.tofile(fr)
    dev = torch.device("cuda:0")
    for disp in ("standard_hdr_pq", "standard_4k"):
        fv = fv_mod.fvvdp(display_name=disp)
        for mode in ("nearest", "bilinear", "bicubic", "area"):
            for tag, res in (("up", (200, 130)), ("down", (116, 75))):
                key = f"{disp}_{mode}_{tag}"
                vs = vy.fvvdp_video_source_yuv_file(ft, fr, display_photometry=disp, full_screen_resize=mode, resize_resolution=res)
                assert list(vs.get_video_size()) == [res[1], res[0], 4]
                rgb0 = vs.test_vidr.get_frame_rgb_tensor(1, dev)
                rgb = vs.test_vidr.get_frame_rgb_tensor(1, dev, vs.resize_of(vs.test_vidr))
                want = torch.nn.functional.interpolate(rgb0.permute(2, 0, 1)[None], size=(res[1], res[0]), mode=mode).clip(0.0, 1.0)[0].permute(1, 2, 0)
                assert tuple(rgb.shape) == (res[1], res[0], 3)
                assert float((rgb - want).abs().max()) < 3e-6, key
                lum = vs.get_test_frame(1, dev)
                assert tuple(lum.shape) == (1, 1, 1, res[1], res[0])
                np.testing.assert_allclose(lum.cpu().numpy()[0, 0, 0], g["lum_" + key], rtol=5e-4, atol=2e-4, err_msg=key)
                if disp == "standard_hdr_pq":
                    jod, st = fv.predict_video_source(vs)  # block path, resize in yuv_resize_planes_kernel
                    assert fv.last_run["h2d_bytes"] == 2 * 4 * t.shape[1] * t.itemsize  # the frames travel at the clip's own size
                    check_jod(jod, g["jod_" + key])
                    # 4 small frames, PQ: the luminance front end's 7e-5 relative differences (fast exp2/log2 in the EOTF) show up
                    # as up to 3e-4 of a channel's largest band energy; the JOD bound above is the stated one (1e-4)
                    check_q(st["Q_per_ch"], g["Q_" + key], tol=5e-4)
    with pytest.raises(ValueError):
        vy.fvvdp_video_source_yuv_file(ft, fr, full_screen_resize="lanczos", resize_resolution=(200, 130)).get_test_frame(0, dev)


def test_batch_front_end(fv_mod, tmp_path, capsys):
    """Pair-level scheduling (two workers on one GPU) returns, in input order, what scoring each pair alone returns; the
    command line prints them and writes the feature files."""
    import cv2
    from fovvideovdp_b200 import run_fvvdp as rf
    from fovvideovdp_b200 import video_source_yuv as vy
    from fovvideovdp_b200.synthetic import synth_yuv_pair
    files = []
    for k, (H, W) in enumerate(((48, 64), (64, 96))):
        t, r = synth_yuv_pair(4 + k, H, W, 10, "420")
        props = dict(width=W, height=H, bit_depth=10, color_space="709", chroma_ss="420", fps=30)
        ft, fr = str(tmp_path / vy.create_yuv_fname(f"t{k}", props)), str(tmp_path / vy.create_yuv_fname(f"r{k}", props))
        t.tofile(ft)
        r.tofile(fr)
        files.append((ft, fr))
    ti, ri = synth_pair_numpy(1, 72, 100)
    cv2.imwrite(str(tmp_path / "t.png"), np.round(ti[0, 0, 0] * 65535).astype(np.uint16))
    cv2.imwrite(str(tmp_path / "r.png"), np.round(ri[0, 0, 0] * 65535).astype(np.uint16))
    files.append((str(tmp_path / "t.png"), str(tmp_path / "r.png")))
    res = rf.score_pairs(files, display="standard_fhd", metrics=("fvvdp", "pu-psnr"), devices=[0, 0])
    assert len(res) == 3
    fv = fv_mod.fvvdp(display_name="standard_fhd")
    for (ft, fr), out in zip(files, res):
        alone, st = fv.predict_video_source(fv_mod.fvvdp_video_source_file(ft, fr, display_photometry="standard_fhd"))
        assert abs(out["FovVideoVDP"][0] - float(alone)) < 1e-6 and np.array_equal(out["FovVideoVDP"][1]["Q_per_ch"], st["Q_per_ch"])
        assert out["PU21-PSNR"][0] > 10
    capsys.readouterr()
    rc = rf.main(["--test"] + [f[0] for f in files] + ["--ref"] + [f[1] for f in files] + ["--display", "standard_fhd", "--quiet", "--features",
                 "--output-dir", str(tmp_path / "out"), "--gpus", "0"])
    printed = [float(v) for v in capsys.readouterr().out.split()]
    assert rc == 0 and len(printed) == 3 and all(abs(a - b["FovVideoVDP"][0]) < 1e-4 for a, b in zip(printed, res))
    assert (tmp_path / "out" / "t_fmap.json").is_file()


def test_pu_psnr(fv_mod, golden):
    """PU21-PSNR (the reference CLI's --metrics pu-psnr): one squared-error kernel per frame pair."""
    g = golden("pu_psnr")
    t, r = synth_pair_numpy(5, 135, 240)
    m = fv_mod.pu_psnr(display_name="standard_4k")
    q, stats = m.predict(t, r, frames_per_second=30)
    assert stats is None and m.short_name() == "PU21-PSNR" and m.quality_unit() == "dB" and m.get_info_string() is None
    assert abs(float(q) - float(g["standard_4k"])) < 2e-3  # dB
    q, _ = fv_mod.pu_psnr(display_name="standard_hdr_pq").predict(0.1 + 0.65 * t, 0.1 + 0.65 * r, frames_per_second=30)
    assert abs(float(q) - float(g["standard_hdr_pq"])) < 2e-3
    q, _ = fv_mod.pu_psnr(display_name="standard_fhd").predict(g["test_u8"], g["ref_u8"], dim_order="FHWC", frames_per_second=30)
    assert abs(float(q) - float(g["u8_rgb_fhd"])) < 2e-3


def test_custom_geometry_foveated(fv_mod, golden):
    """Foveated scoring with a fvvdp_display_geometry SUBCLASS (pytorch_examples/ex_custom_ppd.py:38-57): the per-band
    view-direction / resolution-magnification maps come from the plugin's own methods."""
    class custom_display_geometry(fv_mod.fvvdp_display_geometry):
        def get_ppd(self, view_dir=None):
            if view_dir is None:
                return self.ppd_centre
            view_angle = torch.sqrt(torch.sum((view_dir) ** 2, dim=0, keepdim=False))
            return self.ppd_centre / (view_angle / 20. + 1.)

    g = golden("video_custom_geometry_foveated")
    t, r = synth_pair_numpy(12, 270, 480)
    geo = custom_display_geometry([480, 270], distance_m=0.6, diagonal_size_inches=24)
    fv = fv_mod.fvvdp(display_name="standard_fhd", display_geometry=geo, foveated=True)
    jod, st = fv.predict(t[:, :, :6], r[:, :, :6], frames_per_second=30, fixation_point=g["gaze"])
    check_jod(jod, g["jod"])
    np.testing.assert_allclose(st["rho_band"], g["rho_band"], rtol=1e-6)
    check_q(st["Q_per_ch"], g["Q_per_ch"], tol=1e-3)
    # 60 fps: the same plugin maps through the general kernels
    jod60, _ = fv.predict(t[:, :, :6], r[:, :, :6], frames_per_second=60, fixation_point=g["gaze"])
    assert 0 < float(jod60) < 10


@pytest.mark.parametrize("hw", [(135, 240), (136, 241), (67, 97), (64, 64)])
def test_odd_sizes_with_taps(fv_mod, golden, hw):
    """Odd rows / columns at several pyramid levels, incl. the row-parity quirk of gausspyr_reduce
    (fvvdp_lpyr_dec.py:202) at (135,240) and (136,241)."""
    H, W = hw
    g = golden(f"video_4k_{H}x{W}")
    t, r = synth_pair_numpy(3, H, W)
    fv = fv_mod.fvvdp(display_name="standard_4k")
    fv.debug_taps = True
    jod, st = fv.predict(t, r, frames_per_second=24)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])
    check_taps(fv, g, st["Q_per_ch"].shape[0], 2, int(g["tap_frame"]))


def test_u8_rgb_fhwc(fv_mod, golden):
    g = golden("video_u8_rgb_fhwc")
    jod, st = fv_mod.fvvdp(display_name="standard_fhd").predict(g["test"], g["ref"], dim_order="FHWC", frames_per_second=30)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])
    # the same clip already resident on the GPU (strided channel-last view, no upload)
    tt, rr = torch.from_numpy(g["test"]).cuda(), torch.from_numpy(g["ref"]).cuda()
    jod2, st2 = fv_mod.fvvdp(display_name="standard_fhd").predict(tt, rr, dim_order="FHWC", frames_per_second=30)
    assert float(jod2) == float(jod)


def test_u16_rgb_gamma_bt2020(fv_mod, golden):
    g = golden("image_u16_rgb_gamma_bt2020")
    pm = fv_mod.fvvdp_display_photo_eotf(400, contrast=2000, EOTF="gamma", gamma=2.4, E_ambient=100)
    fv = fv_mod.fvvdp(display_name="standard_4k", display_photometry=pm, color_space="BT.2020")
    jod, st = fv.predict(g["test"], g["ref"], dim_order="HWC")
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])


def test_absolute_and_linear(fv_mod, golden):
    t2, r2 = synth_pair_numpy(3, 64, 64)
    ta, ra = (t2 * 300 + 0.001).astype(np.float32), (r2 * 300 + 0.001).astype(np.float32)
    g = golden("video_absolute")
    fv = fv_mod.fvvdp(display_name="standard_4k", display_photometry=fv_mod.fvvdp_display_photo_absolute(L_max=4000, L_min=0.01))
    jod, st = fv.predict(ta, ra, frames_per_second=30)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])
    g = golden("video_hdr_linear")
    jod, st = fv_mod.fvvdp(display_name="standard_hdr_linear").predict(ta, ra, frames_per_second=30)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])


# ---------------------------------------------------------------------------------------------- oracle, same inputs
def test_against_oracle_random_content(fv_mod, oracle):
    """Seeded noise + structure (not the analytic pattern), uint8 RGB, 60 fps (15 taps)."""
    rng = np.random.default_rng(1234)
    ref = rng.integers(0, 256, size=(6, 90, 121, 3), dtype=np.uint8)
    ref[:, 20:60, 30:90] = (ref[:, 20:60, 30:90] // 4 + 150).astype(np.uint8)
    test = np.clip(ref.astype(np.int32) + rng.integers(-12, 13, size=ref.shape), 0, 255).astype(np.uint8)
    want, wst = oracle.predict(test, ref, dim_order="FHWC", frames_per_second=60, display_name="standard_fhd")
    jod, st = fv_mod.fvvdp(display_name="standard_fhd").predict(test, ref, dim_order="FHWC", frames_per_second=60)
    check_jod(jod, want)
    check_q(st["Q_per_ch"], wst["Q_per_ch"])


def test_against_oracle_1080p_video(fv_mod, oracle):
    """BASELINE config 2 at full frame size, 3 frames (the oracle needs a few seconds per frame)."""
    t, r = synth_pair_numpy(3, 1080, 1920)
    want, wst = oracle.predict(t, r, frames_per_second=30, display_name="standard_fhd")
    jod, st = fv_mod.fvvdp(display_name="standard_fhd").predict(t, r, frames_per_second=30)
    check_jod(jod, want)
    check_q(st["Q_per_ch"], wst["Q_per_ch"])


def test_against_oracle_4k_image(fv_mod, oracle):
    """One full 3840x2160 frame pair (BASELINE config 3 frame size), scored as an image."""
    t, r = synth_pair_numpy(1, 2160, 3840)
    want, wst = oracle.predict(t[0, 0, 0], r[0, 0, 0], dim_order="HW", display_name="standard_4k")
    jod, st = fv_mod.fvvdp(display_name="standard_4k").predict(t[0, 0, 0], r[0, 0, 0], dim_order="HW")
    check_jod(jod, want)
    check_q(st["Q_per_ch"], wst["Q_per_ch"])


def test_against_oracle_4k_foveated_pq_image(fv_mod, oracle):
    """One full 3840x2160 PQ frame pair scored foveated on standard_hdr_pq with an off-centre gaze (the frame size and
    display of BASELINE config 5): view-direction tables, per-pixel rho cells and the re-laid-out CSF table at full size."""
    t, r = synth_pair_numpy(1, 2160, 3840)
    tq, rq = 0.1 + 0.65 * t[0, 0, 0], 0.1 + 0.65 * r[0, 0, 0]
    gaze = np.array([900.0, 1700.0], dtype=np.float32)
    want, wst = oracle.predict(tq, rq, dim_order="HW", display_name="standard_hdr_pq", foveated=True, fixation_point=gaze)
    jod, st = fv_mod.fvvdp(display_name="standard_hdr_pq", foveated=True).predict(tq, rq, dim_order="HW", fixation_point=gaze)
    check_jod(jod, want)
    check_q(st["Q_per_ch"], wst["Q_per_ch"], tol=1e-3)


def test_replicated_first_frame_shortcut_is_exact(fv_mod, monkeypatch):
    """Replicate padding: the kernels reduce the repeated first frame once and copy it into the ring; the result must be
    bit-identical to walking every repeat (FVVDP_B200_NO_DUP_SKIP=1), on the 8- and the 16-frame rings."""
    t, r = synth_pair_torch(20, 270, 480, torch.device("cuda:0"))
    for fps in (30, 60):
        monkeypatch.delenv("FVVDP_B200_NO_DUP_SKIP", raising=False)
        a, sa = fv_mod.fvvdp(display_name="standard_fhd").predict(t, r, frames_per_second=fps)
        monkeypatch.setenv("FVVDP_B200_NO_DUP_SKIP", "1")
        b, sb = fv_mod.fvvdp(display_name="standard_fhd").predict(t, r, frames_per_second=fps)
        assert float(a) == float(b) and np.array_equal(sa["Q_per_ch"], sb["Q_per_ch"])


# ---------------------------------------------------------------------------------------------- plumbing variants
def test_block_size_and_residency_do_not_change_results(fv_mod):
    t, r = synth_pair_numpy(13, 135, 240)
    base, bst = fv_mod.fvvdp(display_name="standard_4k").predict(t, r, frames_per_second=30)
    for bf in (1, 5):
        jod, st = fv_mod.fvvdp(display_name="standard_4k", block_frames=bf).predict(t, r, frames_per_second=30)
        np.testing.assert_allclose(st["Q_per_ch"], bst["Q_per_ch"], rtol=1e-6, atol=1e-9)
        assert abs(float(jod) - float(base)) < 1e-6
    tt, rr = torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda()
    jod, st = fv_mod.fvvdp(display_name="standard_4k", block_frames=5).predict(tt, rr, frames_per_second=30)
    np.testing.assert_allclose(st["Q_per_ch"], bst["Q_per_ch"], rtol=1e-6, atol=1e-9)
    assert isinstance(jod, torch.Tensor) and jod.dim() == 0 and jod.device.type == "cuda"


def test_generic_video_source_matches_array_source(fv_mod):
    """A user-defined fvvdp_video_source (luminance frames from get_*_frame) and a custom photometry subclass go
    through the plug-in path (EOTF by the plug-in, then the kernels) and must agree with the fused path."""
    t, r = synth_pair_numpy(6, 96, 160)
    fv = fv_mod.fvvdp(display_name="standard_fhd", temp_padding="pingpong")
    base, bst = fv.predict(t, r, frames_per_second=30)

    class my_photometry(fv_mod.fvvdp_display_photo_eotf):  # subclass => not restated in-kernel
        def forward(self, V):
            return super().forward(V)

    pm = my_photometry(200, contrast=1000, EOTF="sRGB", E_ambient=250)
    inner = fv_mod.fvvdp_video_source_array(t, r, 30, display_photometry=pm)

    class my_source(fv_mod.fvvdp_video_source):
        def get_video_size(self):
            return inner.get_video_size()

        def get_frames_per_second(self):
            return 30

        def get_test_frame(self, frame, device):
            return inner.get_test_frame(frame, device)

        def get_reference_frame(self, frame, device):
            return inner.get_reference_frame(frame, device)

    jod, st = fv.predict_video_source(my_source())
    check_jod(jod, base, rtol=1e-5)  # torch pow() vs the kernel's exp2/log2 EOTF
    check_q(st["Q_per_ch"], bst["Q_per_ch"], tol=1e-4)
    fv2 = fv_mod.fvvdp(display_name="standard_fhd", display_photometry=pm, temp_padding="pingpong")
    jod, st = fv2.predict(t, r, frames_per_second=30)
    check_jod(jod, base, rtol=1e-5)


def test_errors_and_warnings(fv_mod, caplog):
    fv = fv_mod.fvvdp(display_name="standard_fhd")
    t, r = synth_pair_numpy(2, 64, 64)
    with pytest.raises(RuntimeError):
        fv.predict(t, r)  # video without frames_per_second
    with pytest.raises(RuntimeError):
        fv.predict(t, r[:, :, :1], frames_per_second=30)
    with pytest.raises(RuntimeError):
        fv.predict(t.astype(np.float64), r.astype(np.float64), frames_per_second=30)
    with pytest.raises(RuntimeError):
        fv_mod.fvvdp(display_name="standard_fhd", device="cpu")
    with pytest.raises(RuntimeError):
        fv_mod.fvvdp(display_name="no_such_display")
    with pytest.raises(AssertionError):
        fv_mod.fvvdp(temp_padding="mirror")
    import logging
    with caplog.at_level(logging.WARNING):
        fv.predict(t * 1.5, r, frames_per_second=30)
    assert any("outside the valid range" in rec.message for rec in caplog.records)
    assert fv.get_info_string() == '"FovVideoVDP v1.2.3, 37.84 [pix/deg], Lpeak=200, Lblack=0.5979 [cd/m^2], non-foveated, (standard_fhd)"'


# ---------------------------------------------------------------------------------------------- full-size golden vectors
@pytest.mark.parametrize("case", [("full_fhd_10f", 10, 1080, 1920, "standard_fhd", False), ("full_4k_9f", 9, 2160, 3840, "standard_4k", False),
                                  ("full_4k_hdr_pq_foveated_9f", 9, 2160, 3840, "standard_hdr_pq", True)])
def test_full_size_against_reference(fv_mod, golden, case):
    """BASELINE.json frame sizes (configs[1], [2], [4]) against JOD / Q_per_ch produced by the UNMODIFIED reference on the CPU
    (tools/gen_golden.py full): the first frames of the analytic clip the benchmark scores."""
    name, N, H, W, disp, fov = case
    g = golden(name)
    t, r = synth_pair_numpy(N, H, W)
    kw = {}
    if fov:
        t, r = 0.1 + 0.65 * t, 0.1 + 0.65 * r
        kw["fixation_point"] = g["gaze"]
    jod, st = fv_mod.fvvdp(display_name=disp, foveated=fov).predict(t, r, frames_per_second=30, **kw)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"], tol=1e-3 if fov else 2e-4)
    # the same clip resident on the GPU (TMA-staged level 0)
    jod2, st2 = fv_mod.fvvdp(display_name=disp, foveated=fov).predict(torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda(), frames_per_second=30, **kw)
    check_jod(jod2, g["jod"])
    check_q(st2["Q_per_ch"], g["Q_per_ch"], tol=1e-3 if fov else 2e-4)


# ---------------------------------------------------------------------------------------------- full-size properties
def test_full_size_properties_4k(fv_mod):
    """BASELINE config 3 frame size (3840x2160, standard_4k), 12 frames resident on the GPU:
    identical clips score exactly 10 JOD; a frame block scored on its own reproduces the same per-frame
    pooled energies as the whole clip (what frame sharding relies on); a static clip has constant Q (the temporal
    filters add the ring positions in a fixed order while the weights rotate, so the rounding of a frame depends on its
    index modulo the ring length: exact across block cuts, ~1e-7 of the luminance from frame to frame)."""
    dev = torch.device("cuda:0")
    t, r = synth_pair_torch(12, 2160, 3840, dev)
    fv = fv_mod.fvvdp(display_name="standard_4k", block_frames=4)
    jod, st = fv.predict(r, r, frames_per_second=30)
    assert float(jod) == 10.0 and np.all(st["Q_per_ch"] == 0)
    jod, st = fv.predict(t, r, frames_per_second=30)
    assert 5.0 < float(jod) < 10.0 and np.all(np.isfinite(st["Q_per_ch"]))
    fv12 = fv_mod.fvvdp(display_name="standard_4k", block_frames=12)
    jod12, st12 = fv12.predict(t, r, frames_per_second=30)
    assert np.array_equal(st12["Q_per_ch"], st["Q_per_ch"])  # bit-identical: the ring position follows the frame index in the clip
    stat_t, stat_r = t[:, :, 3:4].expand(1, 1, 9, 2160, 3840), r[:, :, 3:4].expand(1, 1, 9, 2160, 3840)
    jod_s, st_s = fv.predict(stat_t, stat_r, frames_per_second=30)
    q = st_s["Q_per_ch"]
    assert np.abs(q - q[:, :, :1]).max() <= 1e-6 * q.max()  # the transient response of a static clip is itself ~1e-5 of the sustained one


# ---------------------------------------------------------------------------------------------- round 2: more reference fixtures
def check_q_per_band(Q, gQ, tol):
    """Every band and temporal channel against ITS OWN maximum over the frames (a weak band cannot hide behind a strong one)."""
    scale = np.maximum(np.abs(gQ).max(axis=2, keepdims=True), 1e-9)
    err = np.abs(Q - gQ) / scale
    assert err.max() < tol, err.max(axis=2)


def test_bench_clip_64_frames_against_reference(fv_mod, golden):
    """The benchmark's own clip (BASELINE configs[2]: 3840x2160 x 64 frames, standard_4k, 30 fps) against JOD / Q_per_ch of the
    UNMODIFIED reference on the CPU (tools/gen_golden.py round2), resident on the GPU as bench.py scores it."""
    g = golden("full_4k_64f")
    t, r = synth_pair_torch(64, 2160, 3840, torch.device("cuda:0"))
    jod, st = fv_mod.fvvdp(display_name="standard_4k").predict(t, r, frames_per_second=30)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])
    check_q_per_band(st["Q_per_ch"], g["Q_per_ch"], 1e-3)


@pytest.mark.parametrize("name", ["gog_gamma", "gog_srgb"])
def test_gog_photometry(fv_mod, golden, name):
    """fvvdp_display_photo_gog (fvvdp_display_model.py:253-279): gamma branch and the sRGB branch (gamma = -1)."""
    g = golden(f"video_{name}")
    Y_peak, contrast, gamma, E_amb, k_refl = [float(v) for v in g["gog"]]
    dp = fv_mod.fvvdp_display_photo_gog(Y_peak, contrast=contrast, gamma=gamma if gamma > 0 else -1, E_ambient=E_amb, k_refl=k_refl)
    t, r = synth_pair_numpy(8, 270, 480)
    jod, st = fv_mod.fvvdp(display_name="standard_fhd", display_photometry=dp).predict(t, r, frames_per_second=30)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])


@pytest.mark.parametrize("shape", [(9, 135, 240), (40, 270, 480)])
def test_120_fps_against_reference(fv_mod, golden, shape):
    """120 fps: a 30-tap temporal window (fvvdp.py:228), clips shorter and longer than the window.  Default path: the temporal
    filters in a register walk of their own (front_pairs_kernel), the band kernels on two filtered planes per slot."""
    N, H, W = shape
    g = golden(f"video_120fps_{N}x{H}x{W}")
    t, r = synth_pair_numpy(N, H, W)
    jod, st = fv_mod.fvvdp(display_name="standard_fhd").predict(t, r, frames_per_second=120)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])
    # resident clip (TMA-addressable float frames), cut into blocks: the summation order goes by age, so the cut is invisible
    td, rd = torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda()
    jod2, st2 = fv_mod.fvvdp(display_name="standard_fhd").predict(td, rd, frames_per_second=120)
    check_jod(jod2, g["jod"])
    jod3, st3 = fv_mod.fvvdp(display_name="standard_fhd", block_frames=7).predict(td, rd, frames_per_second=120)
    assert np.array_equal(st3["Q_per_ch"], st2["Q_per_ch"])


def test_source_colour_space_overrides_metric(fv_mod, golden):
    """RGB -> luminance weights come from the video SOURCE (video_source.py:87,206), not from the metric's colour space:
    a BT.2020 array source scored by a metric left at sRGB."""
    g = golden("video_source_bt2020_metric_srgb")
    fv = fv_mod.fvvdp(display_name="standard_4k")
    vs = fv_mod.fvvdp_video_source_array(g["test"], g["ref"], 30, dim_order="FHWC", display_photometry=fv.display_photometry, color_space_name="BT.2020")
    jod, st = fv.predict_video_source(vs)
    check_jod(jod, g["jod"])
    check_q(st["Q_per_ch"], g["Q_per_ch"])
    # and it differs from what sRGB weights would give
    jod_srgb, _ = fv.predict(g["test"], g["ref"], dim_order="FHWC", frames_per_second=30)
    assert abs(float(jod_srgb) - float(g["jod"])) > 1e-4


def test_pu_psnr_identical_frames(fv_mod):
    t, _ = synth_pair_numpy(3, 64, 96)
    q, _ = fv_mod.pu_psnr(display_name="standard_4k").predict(t, t, frames_per_second=30)
    assert float(q) == float("inf")


# ---------------------------------------------------------------------------------------------- kernel paths
@pytest.mark.parametrize("levels", ["7", "0"])
def test_kernel_paths_agree_with_reference(fv_mod, golden, monkeypatch, levels):
    """The warp-specialised kernel on every pyramid level (FVVDP_B200_WS_LEVELS=7; by default it runs level 0 only) and the
    fused kernel alone (=0) against the same reference fixtures, incl. non-replicate padding (warm-up walk of the rings),
    short clips, and the 4K frame size; block cuts stay bit-identical on either path."""
    monkeypatch.setenv("FVVDP_B200_WS_LEVELS", levels)
    for pad in ("replicate", "pingpong", "circular"):
        g = golden(f"video_fhd_{pad}")
        t, r = synth_pair_numpy(12, 270, 480)
        jod, st = fv_mod.fvvdp(display_name="standard_fhd", temp_padding=pad).predict(torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda(), frames_per_second=30)
        check_jod(jod, g["jod"])
        check_q(st["Q_per_ch"], g["Q_per_ch"])
    g = golden("video_short_25fps_pingpong")
    t, r = synth_pair_numpy(5, 270, 480)
    jod, st = fv_mod.fvvdp(display_name="standard_fhd", temp_padding="pingpong").predict(torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda(), frames_per_second=25)
    check_jod(jod, g["jod"])
    for pad in ("replicate", "pingpong", "circular"):  # 15-tap windows: the 16x64-tile build of the warp-specialised kernel
        g = golden(f"video_short_60fps_{pad}")
        t, r = synth_pair_numpy(5, 270, 480)
        jod, st = fv_mod.fvvdp(display_name="standard_fhd", temp_padding=pad).predict(torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda(), frames_per_second=60)
        check_jod(jod, g["jod"])
        check_q(st["Q_per_ch"], g["Q_per_ch"])
    g = golden("full_4k_9f")
    t, r = synth_pair_torch(9, 2160, 3840, torch.device("cuda:0"))
    jod, st = fv_mod.fvvdp(display_name="standard_4k").predict(t, r, frames_per_second=30)
    check_jod(jod, g["jod"])
    check_q_per_band(st["Q_per_ch"], g["Q_per_ch"], 1e-3)
    jod4, st4 = fv_mod.fvvdp(display_name="standard_4k", block_frames=4).predict(t, r, frames_per_second=30)
    assert np.array_equal(st4["Q_per_ch"], st["Q_per_ch"])


def test_sharded_clip_equals_single_gpu_nccl(fv_mod):
    """Frame blocks on two GPUs (one process each, NCCL all-reduce of the pooled energies) == the single-GPU result, bit for bit.
    Needs two visible GPUs; the single-GPU test box skips it (tools/check_sharding_nccl.py is the same check under torchrun)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                        "29533", os.path.join(root, "tools", "check_sharding_nccl.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MISMATCH" not in r.stdout


def test_results_are_repeatable_bit_for_bit(fv_mod, monkeypatch):
    """The warp roles of the warp-specialised kernels hand tiles over through mbarriers (which compute-sanitizer's racecheck does
    not model): a missing hand-off would show as run-to-run differences.  Ten runs per configuration (every level on that kernel,
    8- and 15-tap windows, warm-up walk of the rings, several tiles per SM) must give bit-identical per-frame energies."""
    monkeypatch.setenv("FVVDP_B200_WS_LEVELS", "7")
    dev = torch.device("cuda:0")
    t, r = synth_pair_torch(24, 1080, 1920, dev)
    for fps, pad in ((30, "pingpong"), (60, "circular"), (30, "replicate")):
        fv = fv_mod.fvvdp(display_name="standard_fhd", temp_padding=pad, block_frames=9)
        first = None
        for _ in range(10):
            jod, st = fv.predict(t, r, frames_per_second=fps)
            q = st["Q_per_ch"].copy()
            if first is None:
                first = q
            assert np.array_equal(q, first)
