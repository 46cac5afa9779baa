#!/usr/bin/env python3
"""Convert the CIE 1931 observer tables the reference's host uses for spectrum parameters.

Reads (read-only) /root/reference/src/core/cie.rs and emits
pbrt-rust_b200/tables/cie_tables.npz with

  cie_x, cie_y, cie_z   f32[471]   CIE_X / CIE_Y / CIE_Z        (cie.rs:216,338,460)
  cie_lambda            f32[471]   CIE_LAMBDA, 360..830 nm       (cie.rs:582)
  cie_y_integral        f32        CIE_Y_INTEGRAL                (cie.rs:7)

and, from /root/reference/src/materials/metal.rs:13-53, the copper SPD the metal material
defaults to (copper_wavelengths, copper_n, copper_k: f32[56]).

These are standard colorimetric data, needed on the HOST only: `"blackbody L"`, `"spectrum eta"` and
`"xyz Kd"` scene-file parameters are converted to RGB before anything crosses the C ABI
(paramset.rs:163-250, spectrum.rs:129-156).  This script only runs in the build container (the GPU
box has no /root/reference); its output is committed.
"""
import re
from pathlib import Path

import numpy as np

ROOT = Path("/root/reference/src")
OUT = Path(__file__).resolve().parent.parent / "pbrt-rust_b200" / "tables" / "cie_tables.npz"
NUM = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?")


def floats_of(text, name):
    m = re.search(r"%s\s*:\s*\[[^\]]*\]\s*=\s*\[(.*?)\];" % re.escape(name), text, re.S)
    if not m:
        raise SystemExit("array %s not found" % name)
    body = re.sub(r"_f32|_f64", "", m.group(1))
    body = re.sub(r"//[^\n]*", "", body)
    return np.array([float(t) for t in NUM.findall(body)], dtype=np.float64).astype(np.float32)


def main():
    cie = (ROOT / "core" / "cie.rs").read_text()
    out = {k.lower(): floats_of(cie, k) for k in ("CIE_X", "CIE_Y", "CIE_Z", "CIE_LAMBDA")}
    for k, v in out.items():
        assert len(v) == 471, (k, len(v))
    m = re.search(r"CIE_Y_INTEGRAL\s*:\s*Float\s*=\s*([0-9.eE+-]+)", cie)
    out["cie_y_integral"] = np.float32(float(m.group(1)))
    metal = (ROOT / "materials" / "metal.rs").read_text()
    for key, name in (("copper_wavelengths", "COPPER_WAVE_LENGHTS"), ("copper_n", "COPPERN"), ("copper_k", "COPPERK")):
        mm = re.search(r"%s\s*:\s*\[[^\]]*\]\s*=\s*\[(.*?)\];" % name, metal, re.S)
        if mm is None:
            raise SystemExit(name + " not found in metal.rs")
        out[key] = np.array([float(t) for t in NUM.findall(mm.group(1))], dtype=np.float64).astype(np.float32)
    assert len(out["copper_wavelengths"]) == len(out["copper_n"]) == len(out["copper_k"]), [len(out[k]) for k in out]
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes;", len(out["copper_n"]), "copper samples")


if __name__ == "__main__":
    main()
