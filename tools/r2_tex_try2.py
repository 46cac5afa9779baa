"""Scratch: T1 at bench resolution, crop, GPU vs oracle."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
pkg = importlib.import_module("pbrt-rust_b200")
import oracle
for integ_name, md, spp in (("path", 5, 4), ("path", 5, 16), ("volpath", 5, 4)):
    setup = pkg.scenes.textured_scene(xres=1920, yres=1080, spp=spp, maxdepth=md)
    integ = setup.make_integrator(integrator=integ_name)
    nt = integ.n_tiles()
    crop = (nt // 2 - 32, nt // 2 + 32)
    sc = pkg.Scene(setup.flat)
    got, st = sc.render(integ, tile_range=crop)
    sc.close()
    want, ost = oracle.render(setup.flat, integ, tile_range=crop)
    m = want[:, 3] > 0
    a = oracle.film_resolve(got[m], 1.0); b = oracle.film_resolve(want[m], 1.0)
    d = (a.astype(np.float64) - b) ** 2 / (b.astype(np.float64) ** 2 + 1e-2)
    per = d.mean(axis=1)
    idx = np.argsort(per)[::-1][:8]
    print(integ_name, md, spp, "relMSE", d.mean(), "pixels", m.sum(), "rays", st.intersection_tests, ost["intersection_tests"], "top share", per[idx].sum() / per.sum(), "pixels > 1e-2 rel", (per > 1e-2).sum())
