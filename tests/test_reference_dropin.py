"""Drop-in checks against the UNMODIFIED reference package (pyfvvdp, vendored by tools/vendor_reference.py into git-ignored
baseline/_ref/, which travels to the GPU box): the reference's own command line running on this core after install(), the
reference's own display-model and video-source objects handed to our metric, and the reference itself on cuda:0 (TF32 off)
as a second parity anchor on clips no fixture holds.  Skipped when the reference package is not available."""
import os
import sys

import numpy as np
import pytest
import torch

from fovvideovdp_b200.synthetic import synth_pair_numpy, synth_pair_torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def ref():
    import _refimport
    if _refimport.reference_location() is None:
        pytest.skip("reference package not available (tools/vendor_reference.py)")
    return _refimport.import_reference()


@pytest.fixture()
def installed(ref):
    import fovvideovdp_b200 as m
    m.install()
    yield ref
    m.uninstall()


def run_cli(ref_pkg, argv, capsys):
    from pyfvvdp import run_fvvdp
    capsys.readouterr()
    old = sys.argv
    sys.argv = ["fvvdp"] + argv
    try:
        run_fvvdp.main()
    finally:
        sys.argv = old
    return capsys.readouterr().out


def reference_cpu_jod(ref, test, reference, display, **kw):
    cls = getattr(ref, "_reference_classes", {}).get("fvvdp") or sys.modules["pyfvvdp.fvvdp"].fvvdp
    fv = cls(display_name=display, device=torch.device("cpu"), quiet=True)
    with torch.no_grad():
        q, st = fv.predict(test, reference, **kw)
    return float(q), st


def test_reference_cli_on_png_pair(installed, tmp_path, capsys):
    """install(), then pyfvvdp.run_fvvdp.main() (run_fvvdp.py:177-227) on a 16-bit .png pair: the reference's own CLI prints
    the JOD of this core; it must match the unmodified reference scoring the same files on the CPU."""
    import cv2
    import fovvideovdp_b200 as m
    t, r = synth_pair_numpy(1, 270, 480)
    t16 = np.repeat(np.round(t[0, 0, 0, :, :, None] * 65535).astype(np.uint16), 3, axis=2)
    r16 = np.repeat(np.round(r[0, 0, 0, :, :, None] * 65535).astype(np.uint16), 3, axis=2)
    ft, fr = str(tmp_path / "t.png"), str(tmp_path / "r.png")
    cv2.imwrite(ft, t16)
    cv2.imwrite(fr, r16)
    out = run_cli(installed, ["--test", ft, "--ref", fr, "--display", "standard_fhd", "--quiet", "--features", "--output-dir", str(tmp_path / "o")], capsys)
    got = float(out.split()[-1])
    assert installed.fvvdp is m.fvvdp
    want, _ = reference_cpu_jod(installed, t16, r16, "standard_fhd", dim_order="HWC")
    assert abs(got - want) / want < 1e-4, (got, want)
    assert (tmp_path / "o" / "t_fmap.json").is_file()


def test_reference_cli_on_yuv_pair(installed, tmp_path, capsys):
    """The same command line on a raw 10-bit 4:2:0 .yuv pair, both metrics: the JOD equals the fixture the unmodified reference
    produced for these files (tests/golden/yuv_10b_420_2020.npz)."""
    from conftest import load_golden
    from fovvideovdp_b200 import video_source_yuv as vy
    from fovvideovdp_b200.synthetic import synth_yuv_pair
    g = load_golden("yuv_10b_420_2020")
    H, W, bits = int(g["H"]), int(g["W"]), int(g["bits"])
    t, r = synth_yuv_pair(6, H, W, bits, "420")
    props = dict(width=W, height=H, bit_depth=bits, color_space="2020", chroma_ss="420", fps=float(g["fps"]))
    ft, fr = str(tmp_path / vy.create_yuv_fname("test", props)), str(tmp_path / vy.create_yuv_fname("ref", props))
    t.tofile(ft)
    r.tofile(fr)
    out = run_cli(installed, ["--test", ft, "--ref", fr, "--display", str(g["display"]) if "display" in g.files else "standard_hdr_pq",
                              "--metrics", "fvvdp", "pu-psnr"], capsys)
    vals = {line.split("=")[0]: float(line.split("=")[1].split()[0]) for line in out.strip().splitlines() if "=" in line}
    assert abs(vals["FovVideoVDP"] - float(g["jod"])) / float(g["jod"]) < 1e-4, (vals, float(g["jod"]))
    assert vals["PU21-PSNR"] > 10


@pytest.mark.parametrize("kind", ["eotf_srgb", "eotf_pq", "gog", "absolute"])
def test_reference_objects_handed_to_our_metric(ref, kind):
    """The reference's own fvvdp_display_photo_* and fvvdp_video_source_array OBJECTS passed to our fvvdp (recognised by module
    and class name, display_model.py): same JOD as the unmodified reference (CPU) with the same objects."""
    import fovvideovdp_b200 as m
    from pyfvvdp.fvvdp_display_model import fvvdp_display_photo_absolute, fvvdp_display_photo_eotf, fvvdp_display_photo_gog
    from pyfvvdp.video_source import fvvdp_video_source_array
    rng = np.random.default_rng(11)
    t8 = rng.integers(0, 256, (6, 96, 160, 3), dtype=np.uint8)
    r8 = np.clip(t8.astype(np.int32) + rng.integers(-10, 11, t8.shape), 0, 255).astype(np.uint8)
    tv, rv = torch.tensor(t8), torch.tensor(r8)
    if kind == "eotf_srgb":
        dp = fvvdp_display_photo_eotf(250.0, contrast=800, EOTF="sRGB", E_ambient=50)
    elif kind == "eotf_pq":
        dp = fvvdp_display_photo_eotf(1000.0, contrast=100000, EOTF="PQ", E_ambient=10)
    elif kind == "gog":
        dp = fvvdp_display_photo_gog(300.0, contrast=500, gamma=2.4, E_ambient=100)
    else:
        dp = fvvdp_display_photo_absolute(L_max=400.0, L_min=0.5)
        tv, rv = tv.float() * (400.0 / 255.0), rv.float() * (400.0 / 255.0)
    cs = "BT.2020" if kind == "eotf_pq" else "sRGB"
    ref_cls = getattr(ref, "_reference_classes", {}).get("fvvdp") or sys.modules["pyfvvdp.fvvdp"].fvvdp
    fr = ref_cls(display_name="standard_4k", display_photometry=dp, device=torch.device("cpu"), quiet=True)
    with torch.no_grad():
        want, wst = fr.predict_video_source(fvvdp_video_source_array(tv, rv, 30, dim_order="FHWC", display_photometry=dp, color_space_name=cs))
    fo = m.fvvdp(display_name="standard_4k", display_photometry=dp)
    got, gst = fo.predict_video_source(fvvdp_video_source_array(tv, rv, 30, dim_order="FHWC", display_photometry=dp, color_space_name=cs))
    assert abs(float(got) - float(want)) / float(want) < 1e-4, (float(got), float(want))
    scale = np.maximum(np.abs(wst["Q_per_ch"]).max(axis=(0, 2), keepdims=True), 1e-6)
    assert (np.abs(gst["Q_per_ch"] - wst["Q_per_ch"]) / scale).max() < 2e-4


def test_reference_on_cuda_same_clip(ref):
    """The unmodified reference on cuda:0 (TF32 off) and this core on the same resident clip -- content no fixture holds
    (1080p, 24 frames, seeded noise on the analytic pattern): JOD within 1e-4, every band of Q_per_ch within 1e-3 of ITS OWN maximum."""
    import fovvideovdp_b200 as m
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    t, r = synth_pair_torch(24, 1080, 1920, dev)
    gen = torch.Generator(device=dev).manual_seed(3)
    t = (t + 0.03 * torch.randn(t.shape, device=dev, generator=gen)).clamp(0, 1)
    ref_cls = getattr(ref, "_reference_classes", {}).get("fvvdp") or sys.modules["pyfvvdp.fvvdp"].fvvdp
    fr = ref_cls(display_name="standard_fhd", device=dev, quiet=True)
    with torch.no_grad():
        want, wst = fr.predict(t, r, dim_order="BCFHW", frames_per_second=30)
    got, gst = m.fvvdp(display_name="standard_fhd", device=dev).predict(t, r, frames_per_second=30)
    assert abs(float(got) - float(want)) / float(want) < 1e-4, (float(got), float(want))
    wq, gq = wst["Q_per_ch"], gst["Q_per_ch"]
    band_scale = np.maximum(np.abs(wq).max(axis=2, keepdims=True), 1e-9)   # per band AND channel
    assert (np.abs(gq - wq) / band_scale).max() < 1e-3, (np.abs(gq - wq) / band_scale).max(axis=2)
