"""LDR-FLIP (Andersson et al., "FLIP: A Difference Evaluator for Alternating Images", HPG 2020), restated from the
paper for the image-parity tests (north_star check 3).  Inputs are sRGB images in [0, 1], (h, w, 3); the result is
the per-pixel error map in [0, 1].  `tonemap` is the display transform applied to HDR renders before comparison:
clamp to [0, 1], sRGB OETF (what an image viewer shows for an EXR at exposure 0)."""
import numpy as np
from scipy import ndimage

_QC, _QF, _PC, _PT = 0.7, 0.5, 0.4, 0.95
_RGB2XYZ = np.array([[10135552 / 24577794, 8788810 / 24577794, 4435075 / 24577794],
                     [2613072 / 12288897, 8788810 / 12288897, 887015 / 12288897],
                     [1425312 / 73733382, 8788810 / 73733382, 70074185 / 73733382]])
_XYZ2RGB = np.linalg.inv(_RGB2XYZ)
_WHITE = _RGB2XYZ @ np.ones(3)


def tonemap(img):
    x = np.clip(np.nan_to_num(np.asarray(img, dtype=np.float64)), 0.0, 1.0)
    return np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.power(x, 1 / 2.4) - 0.055)


def _srgb_to_linear(s):
    return np.where(s <= 0.04045, s / 12.92, np.power((s + 0.055) / 1.055, 2.4))


def _linrgb_to_ycxcz(rgb):
    xyz = rgb @ _RGB2XYZ.T / _WHITE
    return np.stack([116 * xyz[..., 1] - 16, 500 * (xyz[..., 0] - xyz[..., 1]), 200 * (xyz[..., 1] - xyz[..., 2])], axis=-1)


def _ycxcz_to_linrgb(c):
    y = (c[..., 0] + 16) / 116
    xyz = np.stack([y + c[..., 1] / 500, y, y - c[..., 2] / 200], axis=-1) * _WHITE
    return xyz @ _XYZ2RGB.T


def _linrgb_to_lab(rgb):
    t = rgb @ _RGB2XYZ.T / _WHITE
    d = 6 / 29
    f = np.where(t > d ** 3, np.cbrt(np.maximum(t, 0)), t / (3 * d * d) + 4 / 29)
    return np.stack([116 * f[..., 1] - 16, 500 * (f[..., 0] - f[..., 1]), 200 * (f[..., 1] - f[..., 2])], axis=-1)


def _hunt(lab):
    return np.stack([lab[..., 0], 0.01 * lab[..., 0] * lab[..., 1], 0.01 * lab[..., 0] * lab[..., 2]], axis=-1)


def _hyab(a, b):
    d = a - b
    return np.abs(d[..., 0]) + np.sqrt(d[..., 1] ** 2 + d[..., 2] ** 2)


def _csf_kernels(ppd):
    params = {"A": (1, 0.0047, 0, 1e-5), "RG": (1, 0.0053, 0, 1e-5), "BY": (34.1, 0.04, 13.5, 0.025)}
    r = int(np.ceil(3 * np.sqrt(0.04 / (2 * np.pi ** 2)) * ppd))
    x, y = np.meshgrid(np.arange(-r, r + 1), np.arange(-r, r + 1))
    z = (x / ppd) ** 2 + (y / ppd) ** 2
    out = []
    for key in ("A", "RG", "BY"):
        a1, b1, a2, b2 = params[key]
        g = a1 * np.sqrt(np.pi / b1) * np.exp(-np.pi ** 2 * z / b1) + a2 * np.sqrt(np.pi / b2) * np.exp(-np.pi ** 2 * z / b2)
        out.append(g / g.sum())
    return out


def _feature_kernels(ppd):
    sd = 0.5 * 0.082 * ppd
    r = int(np.ceil(3 * sd))
    x, y = np.meshgrid(np.arange(-r, r + 1), np.arange(-r, r + 1))
    g = np.exp(-(x ** 2 + y ** 2) / (2 * sd * sd))
    out = []
    for gx in (-x * g, (x ** 2 / (sd * sd) - 1) * g):
        neg, pos = -gx[gx < 0].sum(), gx[gx > 0].sum()
        out.append(np.where(gx < 0, gx / neg, gx / pos))
    return out  # edge, point (x direction; y direction = transpose)


def _conv(img, k):
    return ndimage.correlate(img, k, mode="nearest")


def flip_map(reference, test, ppd=67.02):
    ref, tst = np.clip(np.asarray(reference, np.float64), 0, 1), np.clip(np.asarray(test, np.float64), 0, 1)
    yr, yt = _linrgb_to_ycxcz(_srgb_to_linear(ref)), _linrgb_to_ycxcz(_srgb_to_linear(tst))
    # colour pipeline: contrast-sensitivity filtering per opponent channel, then a perceptually uniform distance
    ks = _csf_kernels(ppd)
    fr = np.stack([_conv(yr[..., c], ks[c]) for c in range(3)], axis=-1)
    ft = np.stack([_conv(yt[..., c], ks[c]) for c in range(3)], axis=-1)
    lr = _hunt(_linrgb_to_lab(np.clip(_ycxcz_to_linrgb(fr), 0, 1)))
    lt = _hunt(_linrgb_to_lab(np.clip(_ycxcz_to_linrgb(ft), 0, 1)))
    de = _hyab(lr, lt) ** _QC
    green, blue = _hunt(_linrgb_to_lab(np.array([0.0, 1.0, 0.0]))), _hunt(_linrgb_to_lab(np.array([0.0, 0.0, 1.0])))
    cmax = _hyab(green, blue) ** _QC
    pccmax = _PC * cmax
    dec = np.where(de < pccmax, _PT / pccmax * de, _PT + (de - pccmax) / (cmax - pccmax) * (1 - _PT))
    # feature pipeline: edges and points of the achromatic channel
    ar, at = (yr[..., 0] + 16) / 116, (yt[..., 0] + 16) / 116
    fe = []
    for k in _feature_kernels(ppd):
        nr = np.sqrt(_conv(ar, k) ** 2 + _conv(ar, k.T) ** 2)
        nt = np.sqrt(_conv(at, k) ** 2 + _conv(at, k.T) ** 2)
        fe.append(np.abs(nr - nt))
    def_ = (np.maximum(fe[0], fe[1]) / np.sqrt(2)) ** _QF
    return np.clip(dec, 0, 1) ** (1 - np.clip(def_, 0, 1))


def flip_mean(reference, test, ppd=67.02):
    return float(flip_map(reference, test, ppd).mean())
