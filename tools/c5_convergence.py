"""BASELINE.json configs[4] (C5) as stated: the 5.2 M-triangle glass knot, maxdepth 32, 1024x1024, 2048 spp -- the whole run on one B200,
and the convergence of the estimate: relMSE of the first n samples per pixel against the 2048-spp image, n = 1 .. 1024, plus the relMSE
between the two independent halves (samples [0,1024) vs [1024,2048)) as the noise floor of a 1024-spp image, plus a GPU-vs-oracle crop.
Run under gpurun; writes gpurun_out/c5_convergence.json / .md (kept under profiles/)."""
import importlib, json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
P = importlib.import_module("pbrt-rust_b200")
from oracle import oracle as O
SPP = int(os.environ.get("C5_SPP", "2048"))
setup = P.scenes.glass_knot_scene(nu=4096, nv=640)
flat = setup.flat
integ = setup.make_integrator(spp_=SPP)
film = integ.film
npix = film.width * film.height
sc = P.Scene(flat)
film_t = torch.zeros((npix, 4), dtype=torch.float32, device="cuda")


def render(lo, hi):
    film_t.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, st = sc.render(integ, sample_range=(lo, hi), device_ptr=film_t.data_ptr())
    e1.record(); torch.cuda.synchronize()
    return sc.film_resolve(film_t.cpu().numpy(), film.scale), st, e0.elapsed_time(e1)


render(0, 16)  # warm-up
full, st, ms = render(0, SPP)
rows = {"config": "C5: 5,242,884-triangle glass torus knot, maxdepth 32, 1024x1024, Sobol, %d spp" % SPP, "full_run_ms": ms, "camera_samples": int(st.camera_rays),
        "samples_per_s": st.camera_rays / ms * 1e3, "mrays_per_s": (st.intersection_tests + st.shadow_tests) / ms / 1e3,
        "rays_per_sample": (st.intersection_tests + st.shadow_tests) / st.camera_rays, "iterations": int(st.iterations), "curve": []}
n = 1
while n <= SPP // 2:
    img, s2, m2 = render(0, n)
    rows["curve"].append({"spp": n, "relmse_vs_full": O.rel_mse(img, full), "ms": m2, "samples_per_s": s2.camera_rays / m2 * 1e3})
    print(rows["curve"][-1], flush=True)
    n *= 2
a, _, _ = render(0, SPP // 2)
b, _, _ = render(SPP // 2, SPP)
rows["relmse_between_independent_halves"] = O.rel_mse(a, b)
nt = integ.n_tiles()
crop = (nt // 2, nt // 2 + 64)
g, _ = sc.render(integ, tile_range=crop, sample_range=(0, 8))
c, _ = O.render(flat, integ, nthreads=os.cpu_count() or 1, tile_range=crop, sample_range=(0, 8))
m = c[:, 3] > 0
rows["relmse_gpu_vs_oracle_crop_8spp"] = O.rel_mse(sc.film_resolve(g[m], film.scale), O.film_resolve(c[m], film.scale))
sc.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/c5_convergence.json", "w"), indent=1)
with open("gpurun_out/c5_convergence.md", "w") as f:
    f.write("C5 as BASELINE.json states it (`tools/c5_convergence.py`, one B200): %s\n\n" % rows["config"])
    f.write("* full run: %.1f ms for %d camera samples = **%.1f M samples/s**, %.0f Mrays/s (%.2f rays per sample, %d wavefront iterations)\n" %
            (ms, rows["camera_samples"], rows["samples_per_s"] / 1e6, rows["mrays_per_s"], rows["rays_per_sample"], rows["iterations"]))
    f.write("* relMSE between the two independent %d-spp halves: %.3e (the noise floor of a %d-spp image; the relMSE of an n-spp image falls as 1/n)\n" %
            (SPP // 2, rows["relmse_between_independent_halves"], SPP // 2))
    f.write("* GPU vs CPU oracle on a 64-tile crop at 8 spp: relMSE %.3e (gate: 1e-3)\n\n" % rows["relmse_gpu_vs_oracle_crop_8spp"])
    f.write("| spp (first n samples of every pixel) | relMSE vs the %d-spp image | relMSE x n | ms | M samples/s |\n|---:|---:|---:|---:|---:|\n" % SPP)
    for r in rows["curve"]:
        f.write("| %d | %.3e | %.3e | %.1f | %.1f |\n" % (r["spp"], r["relmse_vs_full"], r["relmse_vs_full"] * r["spp"], r["ms"], r["samples_per_s"] / 1e6))
print(open("gpurun_out/c5_convergence.md").read())
