"""Scratch: textured scenes on the GPU vs the oracle (run on the GPU box)."""
import importlib, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
pkg = importlib.import_module("pbrt-rust_b200")
import oracle


def one(name, flat, integ):
    sc = pkg.Scene(flat)
    t = time.time(); got, st = sc.render(integ); dt = time.time() - t
    sc.close()
    want, ost = oracle.render(flat, integ)
    a = oracle.film_resolve(got, integ.film.scale); b = oracle.film_resolve(want, integ.film.scale)
    print(name, "relMSE %.3e" % oracle.rel_mse(a, b), "mean", a.mean(0), b.mean(0), "finite", np.isfinite(a).all(), "gpu s %.3f" % dt, "dev ms %.2f" % st.device_ms,
          "rays", st.camera_rays, ost["camera_rays"], st.intersection_tests, ost["intersection_tests"], st.shadow_tests, ost["shadow_tests"], flush=True)


S = pkg.scenes
for sampler in ("sobol", "halton", "02sequence"):
    for integ in ("path", "volpath", "whitted", "directlighting:all", "directlighting:one"):
        if integ == "volpath" and sampler == "02sequence":
            continue
        setup = S.textured_scene(xres=160, yres=120, spp=4, sampler=sampler)
        try:
            one(f"T1 {sampler} {integ}", setup.flat, setup.make_integrator(integrator=integ))
        except Exception as e:  # noqa: BLE001
            print(f"T1 {sampler} {integ} FAILED: {e}", flush=True)
setup = S.textured_scene(xres=160, yres=120, spp=4, instanced=False)
one("T1 baked path lens", setup.flat, setup.make_integrator(lensradius=0.05))
one("T1 baked whitted lens gaussian", setup.flat, setup.make_integrator(integrator="whitted", lensradius=0.05, filt="gaussian"))
api = pkg.pbrt_parse(ROOT / "tests" / "golden" / "reference_spheres_scene.pbrt", quick_render=True)
job = api.jobs[0]
one("reference spheres scene (quick)", job.flat, job.integrator)
