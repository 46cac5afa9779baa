"""SURVEY.md §8 f4 measured on one B200 with the last build of round 2 (run under gpurun; writes gpurun_out/r2_f4_sweep.json):
volpath (k_vol_mega) on the fog-box scene at 1024x1024 and on S3 (no media: the same estimator as the path integrator, through the
one-kernel form), whitted / directlighting (k_rec_shade) on S3, the path integrator next to them; each with the CPU oracle's
samples/s on a bounded window of the same render (all host threads) and the image relMSE of a crop against the oracle."""
import importlib, json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
P = importlib.import_module("pbrt-rust_b200")
from oracle import oracle as O
S, H = P.scenes, P.host
nth = os.cpu_count() or 1
rows = []


def measure(name, flat, integ, spp):
    sc = P.Scene(flat)
    film = integ.film
    film_t = torch.zeros((film.width * film.height, 4), dtype=torch.float32, device="cuda")
    sc.render(integ, sample_range=(0, spp), device_ptr=film_t.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    film_t.zero_(); e0.record()
    _, st = sc.render(integ, sample_range=(spp, 2 * spp), device_ptr=film_t.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nt = integ.n_tiles()
    win = (nt // 3, nt // 3 + max(nt // 8, 1))
    t0 = time.time(); _, ost = O.render(flat, integ, nthreads=nth, tile_range=win, sample_range=(0, 4)); t_cpu = time.time() - t0
    crop = (nt // 2, nt // 2 + min(nt // 2, 64))
    a, _ = sc.render(integ, tile_range=crop, sample_range=(0, 4))
    b, _ = O.render(flat, integ, nthreads=nth, tile_range=crop, sample_range=(0, 4))
    m = b[:, 3] > 0
    err = float(O.rel_mse(O.film_resolve(a[m], 1.0), O.film_resolve(b[m], 1.0)))
    sc.close()
    row = {"case": name, "gpu_msamples_per_s": st.camera_rays / ms / 1e3, "gpu_mrays_per_s": (st.intersection_tests + st.shadow_tests) / ms / 1e3, "ms": ms,
           "kernel_launches": int(st.kernel_launches), "cpu_msamples_per_s": ost["camera_rays"] / t_cpu / 1e6, "cpu_threads": nth, "crop_rel_mse": err}
    rows.append(row)
    print(json.dumps(row), flush=True)


fog = S.fog_box_scene(xres=1024, yres=1024, spp=64)
measure("volpath, fog box 1024x1024 (camera in fog, glass with an absorbing medium, material-less boundary), depth 5", fog.flat, fog.make_integrator(spp_=64), 16)
measure("volpath, fog box, depth 40 + roulette", fog.flat, fog.make_integrator(spp_=64, maxdepth_=40), 16)
s3 = S.displaced_sphere_scene()
base = s3.make_integrator(spp_=64)
measure("path, S3 1920x1080 depth 5", s3.flat, base, 16)
measure("volpath, S3 (no media)", s3.flat, H.VolPathIntegrator(base.camera, base.film, base.sampler, maxdepth=5, lightsamplestrategy="power"), 16)
measure("whitted, S3 depth 5", s3.flat, H.WhittedIntegrator(base.camera, base.film, base.sampler, maxdepth=5), 16)
measure("directlighting all, S3 depth 5", s3.flat, H.DirectLightingIntegrator(base.camera, base.film, base.sampler, maxdepth=5, strategy="all"), 16)
measure("directlighting one, S3 depth 5", s3.flat, H.DirectLightingIntegrator(base.camera, base.film, base.sampler, maxdepth=5, strategy="one"), 16)
t1 = S.textured_scene(xres=1920, yres=1080, spp=64)
for integ in ("path", "volpath", "whitted", "directlighting:all"):
    measure(f"{integ}, T1 textured 1920x1080", t1.flat, t1.make_integrator(integrator=integ, spp_=64), 8)
json.dump(rows, open("gpurun_out/r2_f4_sweep.json", "w"), indent=1)
