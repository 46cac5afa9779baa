#!/usr/bin/env python3
"""Freeze the oracle's renders of the sphere-light scene (run from the repo root)."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("pbrt-rust_b200")
from oracle import oracle as O  # noqa: E402

setup = pkg.scenes.sphere_lights_scene()
out = {}
for strategy in ("power", "spatial"):
    out[strategy], _ = O.render_image(setup.flat, setup.make_integrator(spp_=4, res=(48, 32), strategy=strategy))
np.savez_compressed(Path(__file__).parent / "sphere_lights_golden.npz", **out)
print({k: float(v.mean()) for k, v in out.items()})
