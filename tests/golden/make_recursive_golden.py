#!/usr/bin/env python3
"""Freeze the oracle's whitted / directlighting renders of the small mixed scene (run from the repo root)."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
pkg = importlib.import_module("pbrt-rust_b200")
from oracle import oracle as O  # noqa: E402
from test_oracle_recursive_integrators import _integ  # noqa: E402

setup = pkg.scenes.small_mixed_scene()
out = {}
for kind in ("whitted", "one", "all"):
    out[kind], _ = O.render_image(setup.flat, _integ(setup, kind, 4, (48, 32)))
out["all_halton_d3"], _ = O.render_image(setup.flat, _integ(setup, "all", 4, (48, 32), sampler="halton", maxdepth=3))
out["all_zt_d3"], _ = O.render_image(setup.flat, _integ(setup, "all", 4, (48, 32), sampler="02sequence", maxdepth=3))
np.savez_compressed(Path(__file__).parent / "recursive_golden.npz", **out)
print({k: float(v.mean()) for k, v in out.items()})
