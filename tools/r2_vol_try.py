"""Scratch: volpath megakernel vs oracle on the fog box (run on the GPU box)."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
pkg = importlib.import_module("pbrt-rust_b200")
import oracle

def one(name, setup, **kw):
    vol = kw.pop("vol", False)
    integ = setup.make_integrator(**kw)
    if vol:
        integ = pkg.host.VolPathIntegrator(integ.camera, integ.film, integ.sampler, maxdepth=integ.max_depth, rrthreshold=integ.rr_threshold,
                                           lightsamplestrategy=integ.light_sample_strategy)
    sc = pkg.Scene(setup.flat)
    t = time.time(); got, st = sc.render(integ); dt = time.time() - t
    t = time.time(); got, st = sc.render(integ); dt2 = time.time() - t
    sc.close()
    want, ost = oracle.render(setup.flat, integ)
    a = oracle.film_resolve(got, integ.film.scale); b = oracle.film_resolve(want, integ.film.scale)
    print(name, "relMSE %.3e" % oracle.rel_mse(a, b), "mean", a.mean(0), b.mean(0), "gpu s %.3f %.3f" % (dt, dt2), "dev ms %.2f" % st.device_ms,
          "rays", st.camera_rays, ost["camera_rays"], st.intersection_tests, ost["intersection_tests"], "zero", st.zero_radiance_paths, ost["zero_radiance"], flush=True)

S = pkg.scenes
one("fog sobol", S.fog_box_scene())
one("fog halton", S.fog_box_scene(sampler="halton"))
one("fog nocam", S.fog_box_scene(camera_in_fog=False))
one("fog inst", S.fog_box_scene(instanced=True))
one("fog spatial", S.fog_box_scene(), strategy="spatial")
one("fog deep rr", S.fog_box_scene(maxdepth=30), rrthreshold=1.0)
one("cornell vol", S.cornell_scene(xres=128, yres=128, spp=16), vol=True)
