#!/usr/bin/env python
"""Convert the reference's calibration DATA into the package's own data files.

Reads (data only, no code):
  /root/reference/pyfvvdp/csf_cache/o{0,5}_sn1_5_cm0_604562_gpu0.mat   (CSF LUTs, fvvdp.py:505-518)
  /root/reference/pyfvvdp/fvvdp_data/fvvdp_parameters.json             (calibration, fvvdp.py:113-145)
  /root/reference/pyfvvdp/fvvdp_data/display_models.json               (display presets)
  /root/reference/pyfvvdp/fvvdp_data/color_spaces.json                 (RGB->Y weights only)
Writes:
  fovvideovdp_b200/data/csf_lut.npz        float32 axes + S_log[omega][Y][rho][ecc]
  fovvideovdp_b200/data/metric_data.json   {"parameters", "displays", "rgb2y"}

Run once in the build container:  python tools/import_reference_data.py
"""
import json
import os

import numpy as np
import scipy.io

REF = "/root/reference/pyfvvdp"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fovvideovdp_b200", "data")


def load_lut(fname):
    m = scipy.io.loadmat(fname, squeeze_me=True, struct_as_record=False)
    lut = m["lut"]
    return {k: np.asarray(getattr(lut, k), dtype=np.float32) for k in lut._fieldnames}


def main():
    os.makedirs(OUT, exist_ok=True)
    l0 = load_lut(os.path.join(REF, "csf_cache", "o0_sn1_5_cm0_604562_gpu0.mat"))
    l5 = load_lut(os.path.join(REF, "csf_cache", "o5_sn1_5_cm0_604562_gpu0.mat"))
    for k in ("Y", "rho", "ecc", "Y_log", "rho_log", "ecc_sqrt"):
        assert np.array_equal(l0[k], l5[k]), k
    np.savez_compressed(
        os.path.join(OUT, "csf_lut.npz"),
        omega=np.array([0.0, 5.0], dtype=np.float32),
        Y=l0["Y"], rho=l0["rho"], ecc=l0["ecc"],
        Y_log=l0["Y_log"], rho_log=l0["rho_log"], ecc_sqrt=l0["ecc_sqrt"],
        S_log=np.stack([l0["S_log"], l5["S_log"]], 0).astype(np.float32),
    )
    with open(os.path.join(REF, "fvvdp_data", "fvvdp_parameters.json")) as f:
        params = {k: v for k, v in json.load(f).items() if not k.startswith("__")}
    with open(os.path.join(REF, "fvvdp_data", "display_models.json")) as f:
        displays = json.load(f)
    with open(os.path.join(REF, "fvvdp_data", "color_spaces.json")) as f:
        rgb2y = {k: v["RGB2Y"] for k, v in json.load(f).items() if "RGB2Y" in v}
    with open(os.path.join(OUT, "metric_data.json"), "w") as f:
        json.dump({"parameters": params, "displays": displays, "rgb2y": rgb2y}, f, indent=1, sort_keys=True)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
