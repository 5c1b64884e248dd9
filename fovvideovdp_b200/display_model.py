"""Display-model plugin surface: photometry (pixel value -> cd/m^2) and geometry (pixels per degree,
view directions).  Class names, constructor arguments and methods follow the reference's
pyfvvdp/fvvdp_display_model.py so that user code written against it keeps working:

  fvvdp_display_photometry (.forward/.get_peak_luminance/.get_black_level/.print/.load/.list_displays)  :21-98
  fvvdp_display_photo_eotf      :114-189      fvvdp_display_photo_absolute  :191-228
  fvvdp_display_photo_gog       :231-299      fvvdp_display_geometry        :383-568

The metric's CUDA core restates the arithmetic of the stock classes in-kernel (`kernel_spec()` describes an
object to it).  `forward()` / `pix2view_direction()` / `get_resolution_magnification()` remain as torch
code for video sources and custom subclasses, whose results are handed to the kernels as luminance frames
or per-band maps.
"""
import logging
import math

import torch

from . import config

_PQ = dict(Lmax=10000.0, n=0.15930175781250000, m=78.843750000000000, c1=0.83593750000000000, c2=18.851562500000000,
           c3=18.687500000000000)


def srgb2lin(p):
    """sRGB non-linearity (:17-19)."""
    return torch.where(p > 0.04045, ((p + 0.055) / 1.055) ** 2.4, p / 12.92)


def pq2lin(V):
    """SMPTE ST 2084 code values (0-1) -> absolute cd/m^2 (:100-112)."""
    t = torch.pow(V, 1 / _PQ["m"])
    return _PQ["Lmax"] * torch.pow((t - _PQ["c1"]).clamp(min=0) / (_PQ["c2"] - _PQ["c3"] * t), 1 / _PQ["n"])


def _clamp01_with_warning(V):
    if bool(torch.any(V > 1)) or bool(torch.any(V < 0)):
        logging.warning("Pixel outside the valid range 0-1")
        V = V.clamp(0.0, 1.0)
    return V


class fvvdp_display_photometry:
    """Base class of the photometric display models."""

    def forward(self, V):
        raise NotImplementedError

    def print(self):
        raise NotImplementedError

    def kernel_spec(self):
        """Description of this model for the CUDA front end, or None when the model is custom and
        forward() has to be called on every frame."""
        return None

    @classmethod
    def list_displays(cls):
        for name in config.display_models():
            fvvdp_display_photometry.load(name).print()

    @classmethod
    def load(cls, display_name):
        models = config.display_models()
        if display_name not in models:
            raise RuntimeError('Unknown display model: "' + display_name + '"')
        m = models[display_name]
        Y_peak = m["max_luminance"]
        if "min_luminance" in m:
            contrast = Y_peak / m["min_luminance"]
        else:
            contrast = m.get("contrast", 500)
        obj = fvvdp_display_photo_eotf(Y_peak, contrast=contrast, EOTF=m.get("EOTF", "sRGB"), gamma=m.get("gamma", 2.2),
                                       E_ambient=m.get("E_ambient", 0), k_refl=m.get("k_refl", 0.005), name=display_name)
        obj.full_name = m.get("name", display_name)
        obj.short_name = display_name
        return obj


class _reflective_display(fvvdp_display_photometry):
    """Shared by the EOTF and gain-gamma-offset models: peak / contrast / ambient reflection."""

    def get_peak_luminance(self):
        return self.Y_peak

    def get_black_level(self):
        # light reflected from the panel plus the panel's own black level (:172-176)
        return self.E_ambient / math.pi * self.k_refl + self.Y_peak / self.contrast

    def print(self):
        Yb = self.get_black_level()
        logging.info("Photometric display model: {}".format(self.name))
        logging.info("  Peak luminance: {} cd/m^2".format(self.Y_peak))
        if hasattr(self, "EOTF"):
            logging.info("  EOTF: {}".format(self.EOTF))
        logging.info("  Contrast - theoretical: {}:1".format(round(self.contrast)))
        logging.info("  Contrast - effective: {}:1".format(round(self.Y_peak / Yb)))
        logging.info("  Ambient light: {} lux".format(self.E_ambient))
        logging.info("  Display reflectivity: {}%".format(self.k_refl * 100))


class fvvdp_display_photo_eotf(_reflective_display):
    """SDR/HDR display with an 'sRGB', 'gamma', 'PQ' or 'linear' EOTF (:114-189)."""

    def __init__(self, Y_peak, contrast=1000, EOTF="sRGB", gamma=2.2, E_ambient=0, k_refl=0.005, name=None):
        self.Y_peak = Y_peak
        self.contrast = contrast
        self.EOTF = EOTF
        self.gamma = gamma
        self.E_ambient = E_ambient
        self.k_refl = k_refl
        self.name = name

    def forward(self, V):
        if self.EOTF != "linear":
            V = _clamp01_with_warning(V)
        Yb = self.get_black_level()
        if self.EOTF == "sRGB":
            return (self.Y_peak - Yb) * srgb2lin(V) + Yb
        if self.EOTF == "gamma":
            return (self.Y_peak - Yb) * torch.pow(V, self.gamma) + Yb
        if self.EOTF == "PQ":
            return pq2lin(V).clip(0.005, self.Y_peak) + Yb
        if self.EOTF == "linear":
            return V.clip(0.005, self.Y_peak) + Yb
        raise RuntimeError(f"Unknown EOTF '{self.EOTF}'")

    def kernel_spec(self):
        if self.EOTF not in ("sRGB", "gamma", "PQ", "linear"):
            raise RuntimeError(f"Unknown EOTF '{self.EOTF}'")
        return dict(kind=self.EOTF, Y_peak=float(self.Y_peak), Y_black=float(self.get_black_level()), gamma=float(self.gamma))


class fvvdp_display_photo_gog(_reflective_display):
    """Gain-gamma-offset model kept for compatibility; gamma == -1 selects the sRGB curve (:231-299)."""

    def __init__(self, Y_peak, contrast=1000, gamma=2.2, E_ambient=0, k_refl=0.005, name=None):
        self.Y_peak = Y_peak
        self.contrast = contrast
        self.gamma = gamma
        self.E_ambient = E_ambient
        self.k_refl = k_refl
        self.name = name

    def forward(self, V):
        V = _clamp01_with_warning(V)
        Yb = self.get_black_level()
        lin = srgb2lin(V) if self.gamma == -1 else torch.pow(V, self.gamma)
        return (self.Y_peak - Yb) * lin + Yb

    def kernel_spec(self):
        kind = "sRGB" if self.gamma == -1 else "gamma"
        return dict(kind=kind, Y_peak=float(self.Y_peak), Y_black=float(self.get_black_level()), gamma=float(self.gamma))


class fvvdp_display_photo_absolute(fvvdp_display_photometry):
    """Input already is absolute luminance / colorimetric values in cd/m^2 (:191-228)."""

    def __init__(self, L_max=10000, L_min=0.005):
        self.L_max = L_max
        self.L_min = L_min

    def forward(self, V):
        if V.max() < 1:
            logging.warning("Pixel values are very low. Perhaps images are not scaled in the absolute units of cd/m^2.")
        return V.clamp(self.L_min, self.L_max)

    def get_peak_luminance(self):
        return self.L_max

    def get_black_level(self):
        return self.L_min

    def print(self):
        logging.info("Photometric display model:")
        logging.info("  Absolute photometric/colorimetric values")

    def kernel_spec(self):
        return dict(kind="absolute", L_min=float(self.L_min), L_max=float(self.L_max))


# Classes whose arithmetic the CUDA front end restates.  Objects of the reference package itself
# (pyfvvdp.fvvdp_display_model.*) are recognised by module + class name so that a user of the reference can
# pass their existing display objects unchanged.
_STOCK_PHOTOMETRY = {"fvvdp_display_photo_eotf", "fvvdp_display_photo_gog", "fvvdp_display_photo_absolute"}
_STOCK_MODULES = {__name__, "pyfvvdp.fvvdp_display_model"}


def photometry_kernel_spec(pm):
    """kernel_spec() of a stock photometry object (ours or the reference's); None for anything else,
    including subclasses (they may override forward())."""
    t = type(pm)
    if t.__name__ not in _STOCK_PHOTOMETRY or t.__module__ not in _STOCK_MODULES:
        return None
    if t.__module__ == __name__:
        return pm.kernel_spec()
    if t.__name__ == "fvvdp_display_photo_absolute":
        return dict(kind="absolute", L_min=float(pm.L_min), L_max=float(pm.L_max))
    if t.__name__ == "fvvdp_display_photo_gog":
        kind = "sRGB" if pm.gamma == -1 else "gamma"
    else:
        kind = pm.EOTF
        if kind not in ("sRGB", "gamma", "PQ", "linear"):
            raise RuntimeError(f"Unknown EOTF '{kind}'")
    return dict(kind=kind, Y_peak=float(pm.Y_peak), Y_black=float(pm.get_black_level()), gamma=float(pm.gamma))


class fvvdp_display_geometry:
    """Effective resolution in pixels per degree, display size, view directions (:383-568).

    fvvdp_display_geometry(resolution, distance_m=None, distance_display_heights=None, fov_horizontal=None,
                           fov_vertical=None, fov_diagonal=None, diagonal_size_inches=None)
    """

    def __init__(self, resolution, distance_m=None, distance_display_heights=None, fov_horizontal=None, fov_vertical=None,
                 fov_diagonal=None, diagonal_size_inches=None):
        self.resolution = resolution
        self.fixed_ppd = None
        ar = resolution[0] / resolution[1]
        fovs = [fov_horizontal, fov_vertical, fov_diagonal]

        if diagonal_size_inches is not None:
            h_mm = math.sqrt((diagonal_size_inches * 25.4) ** 2 / (1 + ar ** 2))
            self.display_size_m = (ar * h_mm / 1000, h_mm / 1000)

        if distance_m is not None and distance_display_heights is not None:
            raise RuntimeError("You can pass only one of: 'distance_m', 'distance_display_heights'.")
        if distance_m is not None:
            self.distance_m = distance_m
        elif distance_display_heights is not None:
            if not hasattr(self, "display_size_m"):
                raise RuntimeError("You need to specify display diagonal size 'diagonal_size_inches' to specify viewing distance "
                                   "as 'distance_display_heights' ")
            self.distance_m = distance_display_heights * self.display_size_m[1]
        elif any(f is not None for f in fovs):
            self.distance_m = 3  # default for head-mounted displays
        else:
            raise RuntimeError("Viewing distance must be specified as 'distance_m' or 'distance_display_heights'.")

        if sum(f is not None for f in fovs) > 1:
            raise RuntimeError("You can pass only one of 'fov_horizontal', 'fov_vertical', 'fov_diagonal'. The other dimensions "
                               "are inferred from the resolution assuming that the pixels are square.")
        if fov_horizontal is not None:
            w_m = 2 * math.tan(math.radians(fov_horizontal / 2)) * self.distance_m
            self.display_size_m = (w_m, w_m / ar)
        elif fov_vertical is not None:
            h_m = 2 * math.tan(math.radians(fov_vertical / 2)) * self.distance_m
            self.display_size_m = (h_m * ar, h_m)
        elif fov_diagonal is not None:
            # angles do not obey Pythagoras: go through the distance to the screen in pixels
            dist_px = math.hypot(resolution[0], resolution[1]) / (2.0 * math.tan(math.radians(fov_diagonal * 0.5)))
            h_deg = math.degrees(math.atan(resolution[1] / 2 / dist_px)) * 2
            h_m = 2 * math.tan(math.radians(h_deg / 2)) * self.distance_m
            self.display_size_m = (h_m * ar, h_m)
        if not hasattr(self, "display_size_m"):
            raise RuntimeError("Display size must be specified as 'diagonal_size_inches' or one of the 'fov_*' arguments.")

        self.display_size_deg = tuple(2 * math.degrees(math.atan(s / (2 * self.distance_m))) for s in self.display_size_m)
        self.ppd_centre = 1 / (2 * math.degrees(math.atan(0.5 * self.display_size_m[0] / self.resolution[0] / self.distance_m)))

    def get_ppd(self, view_dir=None):
        """Pixels per degree at the screen centre, or for view directions (2,h,w) in degrees (:475-488)."""
        if view_dir is None:
            return self.ppd_centre
        a = torch.sqrt(torch.sum(view_dir ** 2, dim=0))
        a = torch.minimum(a, torch.tensor(89.9, device=a.device))
        delta = (1 / self.ppd_centre) / 2
        tan_a = torch.tan(torch.deg2rad(a))
        return self.ppd_centre * (torch.tan(torch.deg2rad(a + delta)) - tan_a) / math.tan(math.radians(delta))

    def pix2view_direction(self, resolution_pix, x_pix, y_pix):
        """Pixel coordinates (top-left = [0,0]) of an image of `resolution_pix` = [w,h] spanning the display ->
        (2,...) view direction in degrees, x rightwards, y upwards (:498-510)."""
        w, h = float(resolution_pix[0]), float(resolution_pix[1])
        x_m = (x_pix - w / 2) * self.display_size_m[0] / w
        y_m = -(y_pix - h / 2) * self.display_size_m[1] / h
        return torch.stack((torch.rad2deg(torch.atan(x_m / self.distance_m)), torch.rad2deg(torch.atan(y_m / self.distance_m))), dim=0)

    def get_resolution_magnification(self, view_dir):
        """ppd(view_dir) / ppd(centre) (:512-526)."""
        if self.fixed_ppd is not None:
            return torch.ones((), device=view_dir.device)
        return self.get_ppd(view_dir) / self.get_ppd()

    def print(self):
        logging.info("Geometric display model:")
        logging.info("  Resolution: {w} x {h} pixels".format(w=self.resolution[0], h=self.resolution[1]))
        logging.info("  Display size: {w:.1f} x {h:.1f} cm".format(w=self.display_size_m[0] * 100, h=self.display_size_m[1] * 100))
        logging.info("  Display size: {w:.2f} x {h:.2f} deg".format(w=self.display_size_deg[0], h=self.display_size_deg[1]))
        logging.info("  Viewing distance: {d:.3f} m".format(d=self.distance_m))
        logging.info("  Pixels-per-degree (center): {ppd:.2f}".format(ppd=self.get_ppd()))

    @classmethod
    def load(cls, display_name):
        models = config.display_models()
        if display_name not in models:
            raise RuntimeError("Error: Display model '%s' not found in display_models.json" % display_name)
        m = models[display_name]
        assert "resolution" in m
        inch = 0.0254
        if "viewing_distance_meters" in m:
            dist = m["viewing_distance_meters"]
        elif "viewing_distance_inches" in m:
            dist = m["viewing_distance_inches"] * inch
        else:
            dist = None
        if "diagonal_size_meters" in m:
            diag = m["diagonal_size_meters"] / inch
        else:
            diag = m.get("diagonal_size_inches")
        return cls(tuple(m["resolution"]), distance_m=dist, fov_diagonal=m.get("fov_diagonal"), diagonal_size_inches=diag)


_STOCK_GEOMETRY_MODULES = {__name__, "pyfvvdp.fvvdp_display_model"}


def geometry_is_stock(geo):
    """True when the kernels' analytic foveation maps (pix2view_direction / get_ppd of the stock class)
    apply; subclasses (e.g. pytorch_examples/ex_custom_ppd.py:38-57) get per-band maps computed by calling
    their own methods."""
    t = type(geo)
    return t.__name__ == "fvvdp_display_geometry" and t.__module__ in _STOCK_GEOMETRY_MODULES and getattr(geo, "fixed_ppd", None) is None
