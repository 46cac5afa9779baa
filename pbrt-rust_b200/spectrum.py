"""Host-side spectrum parameters -> RGB, as the reference converts them while it parses a scene file.

The reference is built with `Spectrum = RGBSpectrum` (src/core/spectrum.rs:76-160), so every
`"rgb"`, `"xyz"`, `"blackbody"` and `"spectrum"` parameter becomes three f32 coefficients in
`ParamSet::add_*_spectrum` (src/core/paramset.rs:131-250) before any shape, light or material sees it.
Nothing of this runs on the device; the C ABI only ever carries RGB.  All arithmetic is f32 in the
reference's operation order.  Tables: tools/convert_cie_tables.py (CIE 1931 observer, cie.rs).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

f32 = np.float32
_TABLES = None


def cie_tables():
    global _TABLES
    if _TABLES is None:
        t = np.load(Path(__file__).resolve().parent / "tables" / "cie_tables.npz")
        _TABLES = {k: t[k] for k in t.files}
    return _TABLES


def xyz_to_rgb(xyz):  # spectrum.rs:484-492
    x, y, z = (f32(v) for v in xyz)
    return np.array([f32(3.240479) * x - f32(1.537150) * y - f32(0.498535) * z,
                     f32(-0.969256) * x + f32(1.875991) * y + f32(0.041556) * z,
                     f32(0.055648) * x - f32(0.204043) * y + f32(1.057311) * z], f32)


def _interpolate(lam, vals, ls):
    """interpolate_spectrum_samples (spectrum.rs:467-482) at every wavelength of `ls` at once."""
    n = len(lam)
    if n > 1 and not np.all(lam[1:] > lam[:-1]):
        raise ValueError("spectrum samples must have strictly increasing wavelengths (the reference asserts)")
    out = np.empty(len(ls), f32)
    lo, hi = ls <= lam[0], ls >= lam[-1]
    out[lo], out[hi] = vals[0], vals[-1]
    mid = ~(lo | hi)
    if mid.any():
        l = ls[mid]
        # find_interval(n, |i| lambda[i] <= l): last index with lambda <= l, clamped to [0, n-2]
        off = np.clip(np.searchsorted(lam, l, side="right") - 1, 0, n - 2)
        t = (l - lam[off]) / (lam[off + 1] - lam[off])
        out[mid] = (f32(1) - t) * vals[off] + t * vals[off + 1]  # lerp, pbrt.rs
    return out


def from_sampled(lam, vals):
    """RGBSpectrum::from_sampled (spectrum.rs:129-156)."""
    lam, vals = np.asarray(lam, f32), np.asarray(vals, f32)
    if len(lam) > 1 and np.any(lam[:-1] > lam[1:]):
        # spectrum.rs:131-136: sorts (lambda, v) pairs, then recurses with the sorted wavelengths but the
        # UNSORTED values (`&v[..]`, not `&sv[..]`)
        lam = lam[np.lexsort((vals, lam))]
    t = cie_tables()
    val = _interpolate(lam, vals, t["cie_lambda"])
    xyz = np.zeros(3, f32)
    for c, key in enumerate(("cie_x", "cie_y", "cie_z")):
        acc = f32(0)
        for p in (val * t[key]).tolist():  # sequential f32 accumulation, as the loop at :141-146
            acc = f32(acc + f32(p))
        xyz[c] = acc
    n = len(t["cie_lambda"])
    scale = f32(f32(t["cie_lambda"][n - 1]) - f32(t["cie_lambda"][0])) / f32(f32(t["cie_y_integral"]) * f32(n))
    return xyz_to_rgb(xyz * scale)


def black_body(lam_nm, temp):
    """black_body (spectrum.rs:36-58), f32 with the one f64 product the reference has."""
    temp = f32(temp)
    lam_nm = np.asarray(lam_nm, f32)
    if temp <= 0:
        return np.zeros(len(lam_nm), f32)
    c, h, kb = f32(299792458.0), f32(6.62606957e-34), f32(1.3806488e-23)
    with np.errstate(over="ignore"):
        l = (lam_nm.astype(np.float64) * np.float64(f32(1.0e-9))).astype(f32)
        lambda5 = (l * l) * (l * l) * l
        e = np.exp(((h * c) / (l * kb * temp)).astype(f32)).astype(f32)
        return ((f32(2.0) * h * c * c) / (lambda5 * (e - f32(1.0)))).astype(f32)


def black_body_normalized(lam_nm, temp):  # spectrum.rs:60-71
    le = black_body(lam_nm, temp)
    lambda_max = f32(f32(2.8977721e-3) / f32(temp)) * f32(1.0e9)
    return le / black_body(np.array([lambda_max], f32), temp)[0]


def blackbody_rgb(temp, scale):
    """ParamSet::add_blackbody_spectrum (paramset.rs:163-179): normalised Planck curve on the CIE grid x scale."""
    return from_sampled(cie_tables()["cie_lambda"], black_body_normalized(cie_tables()["cie_lambda"], temp)) * f32(scale)


def copper():
    """COPPER_N / COPPER_K of materials/metal.rs:50-53 -> (eta_rgb, k_rgb)."""
    t = cie_tables()
    return from_sampled(t["copper_wavelengths"], t["copper_n"]), from_sampled(t["copper_wavelengths"], t["copper_k"])


def read_float_file(path):
    """read_float_file (src/core/floatfile.rs): whitespace-separated numbers, `#` comments to end of line."""
    vals = []
    with open(path, "r") as f:
        for line in f:
            line = line.split("#", 1)[0]
            for tok in line.split():
                vals.append(float(tok))
    return np.array(vals, f32)
