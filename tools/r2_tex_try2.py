"""Scratch: dump GPU / oracle images for diffing."""
import importlib, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
pkg = importlib.import_module("pbrt-rust_b200")
import oracle
S = pkg.scenes
out = {}
def run(name, flat, it):
    sc = pkg.Scene(flat)
    got, st = sc.render(it)
    sc.close()
    want, ost = oracle.render(flat, it)
    a = oracle.film_resolve(got, 1.0); b = oracle.film_resolve(want, 1.0)
    print(name, oracle.rel_mse(a, b), a.mean(0), b.mean(0), st.intersection_tests, ost["intersection_tests"], st.shadow_tests, ost["shadow_tests"])
    out[name + "_gpu"] = a; out[name + "_cpu"] = b
setup = S.textured_scene(xres=160, yres=120, spp=4, sampler="02sequence")
run("ztpath", setup.flat, setup.make_integrator(integrator="path"))
api = pkg.pbrt_parse(ROOT / "tests" / "golden" / "reference_spheres_scene.pbrt", quick_render=True)
job = api.jobs[0]
run("refscene", job.flat, job.integrator)
print(job.film.width, job.film.height)
np.savez_compressed(ROOT / "gpurun_out" / "tex_imgs.npz", **out)
