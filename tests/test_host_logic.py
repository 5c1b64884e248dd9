"""CPU-only checks: host-side logic of the product package against the golden vectors / the oracle, and the
C-ABI library (loads, exports every symbol include/fvvdp_b200.h declares; no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from fovvideovdp_b200 import _native, config
from fovvideovdp_b200 import build as native_build
from fovvideovdp_b200.display_model import (fvvdp_display_geometry, fvvdp_display_photo_absolute, fvvdp_display_photo_eotf,
                                            fvvdp_display_photo_gog, fvvdp_display_photometry, geometry_is_stock, photometry_kernel_spec)
from fovvideovdp_b200.fvvdp import HALO_SLOT_COST, frame_block, initial_window, pyramid_layout, temporal_filters
from fovvideovdp_b200.video_source import fvvdp_video_source_array, reshuffle_dims
from oracle import fvvdp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_the_declared_abi():
    path = native_build.build_native()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "fvvdp_b200.h")).read()
    declared = set(re.findall(r"\b(fvvdp_b200_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.fvvdp_b200_abi_version() == _native.ABI_VERSION


def test_config_struct_matches_header_size():
    """sizeof(fvvdp_b200_config) as laid out by ctypes == as laid out by the C compiler."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "fvvdp_b200.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(fvvdp_b200_config), sizeof(fvvdp_b200_pool_params), sizeof(fvvdp_b200_yuv_desc), sizeof(fvvdp_b200_pu_params));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "s")])
        a, b, c_, d_ = subprocess.check_output([os.path.join(d, "s")]).decode().split()
    assert int(a) == ctypes.sizeof(_native.Config)
    assert int(b) == ctypes.sizeof(_native.PoolParams)
    assert int(c_) == ctypes.sizeof(_native.YuvDesc)
    assert int(d_) == ctypes.sizeof(_native.PuParams)


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _native.load_library()
    cfg = _native.Config()
    cfg.abi_version = _native.ABI_VERSION
    cfg.width, cfg.height, cfg.n_levels, cfg.temp_ch, cfg.filter_len, cfg.in_channels, cfg.max_block_frames = 64, 64, 3, 1, 1, 1, 1
    lut = config.csf_lut()
    keep = [np.ascontiguousarray(lut[k]) for k in ("rho_log", "Y_log", "ecc_sqrt", "S_log")]
    cfg.csf_rho_log, cfg.csf_Y_log, cfg.csf_ecc_sqrt, cfg.csf_S_log = [a.ctypes.data_as(ctypes.c_void_p) for a in keep]
    h = ctypes.c_void_p()
    rc = lib.fvvdp_b200_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc < 0 and not h.value
    assert b"no CUDA device" in lib.fvvdp_b200_last_error(None)
    import fovvideovdp_b200
    with pytest.raises(RuntimeError):
        fovvideovdp_b200.fvvdp()


def test_pyramid_layout(golden):
    for row in golden("unit_pyr_layout")["rows"]:
        W, H, ppd, height = int(row[0]), int(row[1]), row[2], int(row[3])
        n_levels, f = pyramid_layout(W, H, ppd)
        assert n_levels == height + 1
        np.testing.assert_allclose(f, row[4:4 + height + 1], rtol=1e-12)


def test_temporal_filters(golden):
    g = golden("unit_temporal_filters")
    for fps in (24, 25, 30, 50, 60, 120, 12.5):
        fl = int(np.ceil(250.0 / (1000.0 / fps)))
        np.testing.assert_allclose(temporal_filters(fps, fl, 0.5, 0.06), g[f"F_{fps}"], rtol=1e-4, atol=2e-7)


@pytest.mark.parametrize("pad", ["replicate", "circular", "pingpong"])
def test_window_rule_matches_oracle(pad):
    for N, fl in ((12, 8), (5, 7), (5, 15), (3, 30), (2, 6), (64, 8)):
        first = initial_window(N, fl, pad)
        assert len(first) == fl
        for ff in range(N):
            want = O.window_indices(ff, N, fl, pad)
            got = [(t if t >= 1 else first[fl - 1 + t]) for t in range(ff - fl + 1, ff + 1)]
            assert got == want


def test_frame_block_work_balanced():
    """With a temporal halo the cut balances work, not frames: contiguous, complete, non-empty, and the first rank (whose halo is
    one repeated frame under replicate padding) takes the most frames."""
    for N in (8, 9, 37, 128, 256, 512):
        for G in (2, 4, 8):
            if N < G:
                continue
            blocks = [frame_block(N, r, G, halo=7, first_halo=1) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == N
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(G - 1)) and all(b > a for a, b in blocks)
            sizes = [b - a for a, b in blocks]
            assert sizes[0] == max(sizes) and max(sizes) - min(sizes) <= 7
            work = [sz + HALO_SLOT_COST * (1 if r == 0 else 7) for r, sz in enumerate(sizes)]
            if N >= 16 * G:
                assert max(work) - min(work) <= 1.5


def test_frame_block_partition():
    for N in (1, 7, 64, 255, 256):
        for G in (1, 2, 3, 8):
            blocks = [frame_block(N, r, G) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == N
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_presets_match_reference(golden):
    g = golden("unit_presets")
    for name, row in zip(g["names"], g["rows"]):
        ph, ge = fvvdp_display_photometry.load(str(name)), fvvdp_display_geometry.load(str(name))
        got = [ph.get_peak_luminance(), ph.get_black_level(), ge.get_ppd(), ge.display_size_m[0], ge.display_size_m[1], ge.distance_m]
        np.testing.assert_allclose(got, row, rtol=1e-12)


def test_photometry_forward_matches_reference(golden):
    g = golden("unit_eotf")
    V = torch.from_numpy(g["V"])
    for kind in ("sRGB", "gamma", "PQ", "linear"):
        Yp = 1500 if kind in ("PQ", "linear") else 200
        pm = fvvdp_display_photo_eotf(Yp, contrast=1000, EOTF=kind, gamma=2.2, E_ambient=250)
        Vin = V * 2000 if kind == "linear" else V
        np.testing.assert_allclose(pm.forward(Vin).numpy(), g[kind], rtol=2e-5, atol=1e-6)
        spec = photometry_kernel_spec(pm)
        assert spec["kind"] == kind and abs(spec["Y_black"] - float(g[kind + "_black"])) < 1e-9
    pa = fvvdp_display_photo_absolute(L_max=1000, L_min=0.01)
    np.testing.assert_allclose(pa.forward(V * 2000).numpy(), g["absolute"], rtol=1e-7)
    assert photometry_kernel_spec(pa) == dict(kind="absolute", L_min=0.01, L_max=1000.0)
    assert photometry_kernel_spec(fvvdp_display_photo_gog(100, gamma=-1))["kind"] == "sRGB"

    class custom(fvvdp_display_photo_eotf):
        pass

    assert photometry_kernel_spec(custom(100)) is None  # subclasses go through forward()


def test_geometry_matches_reference(golden):
    g = golden("unit_foveation")
    geo = fvvdp_display_geometry.load("standard_hmd")
    assert geometry_is_stock(geo)
    w, h = [int(v) for v in g["band_wh"]]
    fw, fh = [int(v) for v in g["frame_wh"]]
    xv = torch.linspace(0.5, w - 0.5, w)
    yv = torch.linspace(0.5, h - 0.5, h)
    xx, yy = torch.meshgrid(xv, yv, indexing="xy")
    vd = geo.pix2view_direction(torch.tensor((w, h)), xx, yy)
    gz = geo.pix2view_direction(torch.tensor((fw, fh)), torch.as_tensor(g["gaze"][0] + 0.5), torch.as_tensor(g["gaze"][1] + 0.5)).view(2, 1, 1)
    ecc = torch.sqrt(torch.sum((vd - gz) ** 2, dim=0))
    np.testing.assert_allclose(ecc.numpy(), g["ecc"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(geo.get_resolution_magnification(vd).numpy(), g["res_mag"], rtol=5e-3)
    with pytest.raises(RuntimeError):
        fvvdp_display_geometry((1920, 1080), diagonal_size_inches=24)
    with pytest.raises(RuntimeError):
        fvvdp_display_geometry((1920, 1080), distance_m=1, distance_display_heights=2, diagonal_size_inches=24)
    g2 = fvvdp_display_geometry((1920, 1080), distance_display_heights=3, diagonal_size_inches=47)
    assert abs(g2.distance_m - 3 * g2.display_size_m[1]) < 1e-12

    class custom(fvvdp_display_geometry):
        pass

    assert not geometry_is_stock(custom((1920, 1080), distance_m=1, diagonal_size_inches=24))


def test_array_source_matches_oracle_luminance(golden):
    g = golden("video_u8_rgb_fhwc")
    vs = fvvdp_video_source_array(g["test"], g["ref"], 30, dim_order="FHWC", display_photometry="standard_fhd")
    assert vs.get_video_size() == (72, 50, 3)
    L = vs.get_test_frame(1)[0, 0, 0].numpy()
    photo = O.photometry_from_preset("standard_fhd")
    want = O.frame_luminance(np.transpose(g["test"][1], (2, 0, 1)), photo, O.metric_data()["rgb2y"]["sRGB"])
    np.testing.assert_allclose(L, want, rtol=2e-5)
    u16 = (g["test"].astype(np.uint16) * 257)
    vs16 = fvvdp_video_source_array(u16, u16, 30, dim_order="FHWC", display_photometry="standard_fhd")
    assert vs16.test_video.dtype == torch.int16
    np.testing.assert_allclose(vs16.get_test_frame(1)[0, 0, 0].numpy(), want, rtol=2e-5)
    # shard view: frames [1,3) of a 3-frame clip
    sh = fvvdp_video_source_array(g["test"][1:], g["ref"][1:], 30, dim_order="FHWC", display_photometry="standard_fhd", first_frame=1, total_frames=3)
    assert sh.get_video_size() == (72, 50, 3)
    np.testing.assert_allclose(sh.get_test_frame(1)[0, 0, 0].numpy(), L)
    with pytest.raises(RuntimeError):
        sh.get_test_frame(0)
    with pytest.raises(RuntimeError):
        fvvdp_video_source_array(g["test"], g["ref"], 0, dim_order="FHWC")
    with pytest.raises(RuntimeError):
        fvvdp_video_source_array(g["test"], g["ref"][:2], 30, dim_order="FHWC")
    with pytest.raises(RuntimeError):
        fvvdp_video_source_array(g["test"][..., :2], g["ref"][..., :2], 30, dim_order="FHWC")


def test_reshuffle_dims():
    x = torch.arange(2 * 3 * 4).reshape(2, 3, 4)
    y = reshuffle_dims(x, "HWC", "BCFHW")
    assert tuple(y.shape) == (1, 4, 1, 2, 3)
    assert y[0, 1, 0, 1, 2] == x[1, 2, 1]
    with pytest.raises(RuntimeError):
        reshuffle_dims(x, "HWZ", "BCFHW")


def test_config_search_order(tmp_path, monkeypatch):
    import json
    models = {"my_display": {"name": "x", "resolution": [100, 50], "viewing_distance_meters": 1, "diagonal_size_inches": 10, "max_luminance": 123}}
    (tmp_path / "display_models.json").write_text(json.dumps(models))
    monkeypatch.setenv("FVVDP_PATH", str(tmp_path))
    assert fvvdp_display_photometry.load("my_display").get_peak_luminance() == 123
    monkeypatch.delenv("FVVDP_PATH")
    with pytest.raises(RuntimeError):
        fvvdp_display_photometry.load("my_display")
    assert config.parameters()["version"] == "1.2.3"
    with pytest.raises(RuntimeError):
        config.rgb2y("no such space")


def test_yuv_file_name_properties(tmp_path):
    """decode_video_props / create_yuv_fname / YUVReader geometry (video_source_yuv.py:6-110) -- host logic only."""
    from fovvideovdp_b200 import video_source_yuv as vy
    from fovvideovdp_b200.synthetic import synth_yuv_pair
    props = dict(width=64, height=48, bit_depth=10, color_space="2020", chroma_ss="420", fps=29.97)
    name = vy.create_yuv_fname("clip", props)
    assert name == "clip_64x48_10b_420_2020_29.97fps.yuv"
    assert vy.decode_video_props("/some/dir/" + name) == props
    assert vy.decode_video_props("x_1280x720p_8b_444_bt709_25fps.yuv") == dict(width=1280, height=720, bit_depth=8, color_space="709",
                                                                                  chroma_ss="444", fps=25.0)
    d = vy.decode_video_props("noprops.yuv")  # the reference's defaults (:8-14)
    assert (d["width"], d["height"], d["fps"], d["bit_depth"], d["color_space"], d["chroma_ss"]) == (1920, 1080, 24, 8, "2020", "420")
    t, r = synth_yuv_pair(3, 48, 64, 10, "420")
    assert t.dtype == np.uint16 and t.shape == (3, 48 * 64 * 3 // 2) and t.min() >= 64 and t.max() <= 940
    path = tmp_path / name
    t.tofile(path)
    rd = vy.YUVReader(str(path))
    assert (rd.width, rd.height, rd.frame_count, rd.bit_depth, rd.uv_shape) == (64, 48, 3, 10, (24, 32))
    Y, u, v = rd.get_frame_yuv(2)
    assert Y.shape == (48, 64) and u.shape == (24, 32) and np.array_equal(Y.ravel(), t[2, :48 * 64])
    with pytest.raises(RuntimeError):
        rd.get_frame_yuv(3)
    with pytest.raises(FileNotFoundError):
        vy.YUVReader(str(tmp_path / "missing_64x48.yuv"))
    with pytest.raises(RuntimeError):  # conversion needs the CUDA kernel: no CPU fallback
        rd.get_frame_rgb_tensor(0, torch.device("cpu"))


def test_oracle_chroma_upsampling_is_torch_bilinear():
    """The oracle's 4:2:0 chroma upsampling restates torch.nn.functional.interpolate(scale_factor=2, mode='bilinear')
    (video_source_yuv.py:219-221)."""
    from oracle import fvvdp_oracle as O
    rng = np.random.default_rng(3)
    for shape in ((5, 7), (24, 32), (1, 9)):
        c = rng.random(shape, dtype=np.float32) - 0.5
        want = torch.nn.functional.interpolate(torch.tensor(c)[None, None], scale_factor=2, mode="bilinear")[0, 0].numpy()
        np.testing.assert_allclose(O._upsample2_bilinear(c), want, atol=1e-6)


def test_heatmap_and_geometry_arguments_are_validated():
    import fovvideovdp_b200 as m
    with pytest.raises(AssertionError):
        m.fvvdp.__init__(m.fvvdp.__new__(m.fvvdp), heatmap="rainbow")
    with pytest.raises(AssertionError):
        m.fvvdp.__init__(m.fvvdp.__new__(m.fvvdp), temp_padding="mirror")
    from fovvideovdp_b200.display_model import fvvdp_display_geometry, geometry_is_stock

    class custom(fvvdp_display_geometry):
        def get_ppd(self, view_dir=None):
            return self.ppd_centre if view_dir is None else self.ppd_centre / (torch.sqrt(torch.sum(view_dir ** 2, dim=0)) / 20.0 + 1.0)

    stock = fvvdp_display_geometry([480, 270], distance_m=0.6, diagonal_size_inches=24)
    mine = custom([480, 270], distance_m=0.6, diagonal_size_inches=24)
    assert geometry_is_stock(stock) and not geometry_is_stock(mine)
    view = mine.pix2view_direction(torch.tensor((4, 2)), torch.tensor([[0.5, 3.5]]), torch.tensor([[0.5, 1.5]]))
    assert view.shape == (2, 1, 2) and float(mine.get_resolution_magnification(view).max()) < 1.0


def test_batch_front_end_host_logic(tmp_path):
    """Pair expansion / dealing of the batch front end and the file-type dispatch of fvvdp_video_source_file
    (run_fvvdp.py:156-167,201-212; video_source_file.py:413-443)."""
    import cv2
    from fovvideovdp_b200 import run_fvvdp as rf
    from fovvideovdp_b200.video_source_file import fvvdp_video_source_file, load_image_as_array
    assert rf.deal_pairs(7, 3) == [[0, 3, 6], [1, 4], [2, 5]] and rf.deal_pairs(2, 4) == [[0], [1], [], []]
    assert rf.expand_pairs(["a", "b", "c"], ["r"]) == [("a", "r"), ("b", "r"), ("c", "r")]
    assert rf.expand_pairs(["a"], ["r", "s"]) == [("a", "r"), ("a", "s")]
    with pytest.raises(RuntimeError):
        rf.expand_pairs(["a", "b"], ["r", "s", "t"])
    rng = np.random.default_rng(5)
    rgb16 = rng.integers(0, 65536, (20, 32, 3), dtype=np.uint16)
    grey8 = rng.integers(0, 256, (20, 32), dtype=np.uint8)
    cv2.imwrite(str(tmp_path / "a.png"), rgb16[:, :, ::-1])
    cv2.imwrite(str(tmp_path / "b.png"), np.dstack([rgb16[:, :, ::-1], np.full((20, 32), 65535, np.uint16)]))  # with alpha
    cv2.imwrite(str(tmp_path / "g.png"), grey8)
    assert np.array_equal(load_image_as_array(str(tmp_path / "a.png")), rgb16)
    assert np.array_equal(load_image_as_array(str(tmp_path / "b.png")), rgb16)
    assert load_image_as_array(str(tmp_path / "g.png")).shape == (20, 32, 1)
    vs = fvvdp_video_source_file(str(tmp_path / "a.png"), str(tmp_path / "b.png"), display_photometry="standard_4k")
    assert tuple(vs.get_video_size()) == (20, 32, 1) and vs.get_frames_per_second() == 0
    (tmp_path / "clip_32x20_8b_420_709_25fps.yuv").write_bytes(bytes(32 * 20 * 3 // 2 * 2))
    (tmp_path / "clip.mp4").write_bytes(b"not a video")
    with pytest.raises(AssertionError):
        fvvdp_video_source_file(str(tmp_path / "a.png"), str(tmp_path / "clip_32x20_8b_420_709_25fps.yuv"))
    with pytest.raises(RuntimeError):
        fvvdp_video_source_file(str(tmp_path / "clip.mp4"), str(tmp_path / "clip.mp4"))
    y = fvvdp_video_source_file(str(tmp_path / "clip_32x20_8b_420_709_25fps.yuv"), str(tmp_path / "clip_32x20_8b_420_709_25fps.yuv"),
                                display_photometry="standard_4k")
    assert list(y.get_video_size()) == [20, 32, 2] and y.get_frames_per_second() == 25.0
