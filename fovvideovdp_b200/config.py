"""Configuration / calibration data lookup.

Mirrors the search order of the reference's `utils.config_files` (pyfvvdp/utils.py:129-154): an explicit
configuration directory, then $FVVDP_PATH, then the data shipped with the package.  User directories hold
the reference's own file names (display_models.json, fvvdp_parameters.json, color_spaces.json); the packaged
copy is the merged fovvideovdp_b200/data/metric_data.json + csf_lut.npz written by
tools/import_reference_data.py.
"""
import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class config_files:
    fvvdp_config_dir = None

    @classmethod
    def set_config_dir(cls, path):
        cls.fvvdp_config_dir = path

    @classmethod
    def find(cls, fname):
        """Path of a user-supplied configuration file, or None when only the packaged data applies."""
        for d in (cls.fvvdp_config_dir, os.getenv("FVVDP_PATH")):
            if d is not None:
                p = os.path.join(d, fname)
                if os.path.isfile(p):
                    return p
        return None


_packaged = {}
# calibration the packaged CSF tables were computed for (csf_cache/o{0,5}_sn1_5_cm0_604562_gpu0.mat, tools/import_reference_data.py)
CSF_LUT_KEY = {"csf_sigma": -1.5, "k_cm": 0.604562}


def _packaged_data():
    if "md" not in _packaged:
        with open(os.path.join(_DATA, "metric_data.json")) as f:
            _packaged["md"] = json.load(f)
    return _packaged["md"]


def _user_json(fname):
    p = config_files.find(fname)
    if p is None:
        return None
    with open(p) as f:
        return json.load(f)


def display_models():
    d = _user_json("display_models.json")
    return d if d is not None else _packaged_data()["displays"]


def parameters():
    d = _user_json("fvvdp_parameters.json")
    if d is not None:
        return {k: v for k, v in d.items() if not k.startswith("__")}
    return _packaged_data()["parameters"]


def rgb2y(color_space_name):
    d = _user_json("color_spaces.json")
    table = {k: v["RGB2Y"] for k, v in d.items() if "RGB2Y" in v} if d is not None else _packaged_data()["rgb2y"]
    if color_space_name not in table:
        raise RuntimeError('Unknown color space: "' + color_space_name + '"')
    return [float(v) for v in table[color_space_name]]


def csf_lut():
    """CSF look-up tables (fvvdp.py:505-518): axes (32,) and S_log[omega in {0,5} Hz][Y][rho][ecc], float32."""
    if "lut" not in _packaged:
        d = np.load(os.path.join(_DATA, "csf_lut.npz"))
        _packaged["lut"] = {k: np.ascontiguousarray(d[k], dtype=np.float32) for k in d.files}
    return _packaged["lut"]
