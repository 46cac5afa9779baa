"""Every BASELINE.json config on one B200: device-resident samples/s and Mrays/s, the CPU oracle's samples/s on the same
scene (bounded sample), and the image relMSE between the two on a low-spp crop of the same frame.  Run under gpurun;
writes gpurun_out/config_sweep.json (summarised in profiles/)."""
import importlib, json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
P = importlib.import_module("pbrt-rust_b200")
from oracle import oracle as O
S = P.scenes
only = sys.argv[1:]  # optional subset of config names
CONFIGS = [
    ("C1 spheres 400x400 64spp sobol d5", lambda: S.spheres_scene(), dict(), 64),
    ("C1 spheres 400x400 64spp 02sequence (tile-serial)", lambda: S.spheres_scene(), dict(sampler_="02sequence"), 64),
    ("C2 cornell 1024x1024 d8 gaussian power", lambda: S.cornell_scene(), dict(), 32),
    ("C2 cornell 1024x1024 d8 gaussian spatial (reference default)", lambda: S.cornell_scene(), dict(strategy="spatial"), 32),
    ("C3 S3 1M-tri sphere 1920x1080 d5", lambda: S.displaced_sphere_scene(), dict(), 16),
    ("C3 S3 halton", lambda: S.displaced_sphere_scene(), dict(sampler_="halton"), 16),
    ("C4 S4 20M instanced tris, 10k lights, 3840x2160 d5", lambda: S.foliage_field_scene(), dict(), 4),
    ("C5 S5 5.2M-tri glass knot 1024x1024 d32", lambda: S.glass_knot_scene(nu=4096, nv=640), dict(), 32),
]
out = []
nth = os.cpu_count() or 1
for name, make, kw, spp in CONFIGS:
    if only and not any(o in name for o in only):
        continue
    t0 = time.time(); setup = make(); t_build = time.time() - t0
    flat = setup.flat
    integ = setup.make_integrator(spp_=spp * 3, **kw)
    film = integ.film
    sc = P.Scene(flat)
    film_t = torch.zeros((film.width * film.height, 4), dtype=torch.float32, device="cuda")
    sc.render(integ, sample_range=(0, spp), device_ptr=film_t.data_ptr())  # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    film_t.zero_(); e0.record()
    _, st = sc.render(integ, sample_range=(spp, 2 * spp) if "02sequence" not in name else (0, spp), device_ptr=film_t.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rays = st.intersection_tests + st.shadow_tests
    # CPU oracle: whole frame at 1 spp (or a tile window for the heavy ones), then relMSE on a crop at low spp
    nt = integ.n_tiles()
    win = (0, nt) if film.width * film.height <= 1100 * 1100 else (nt // 3, nt // 3 + max(nt // 12, 1))
    t0 = time.time(); _, ost = O.render(flat, integ, nthreads=nth, tile_range=win, sample_range=(0, 2)); t_cpu = time.time() - t0
    cpu_sps = ost["camera_rays"] / t_cpu
    crop = (nt // 2, nt // 2 + min(nt // 2, 96))
    a, _ = sc.render(integ, tile_range=crop, sample_range=(0, 4))
    b, _ = O.render(flat, integ, nthreads=nth, tile_range=crop, sample_range=(0, 4))
    m = b[:, 3] > 0
    ia, ib = sc.film_resolve(a[m], film.scale), O.film_resolve(b[m], film.scale)
    rel = O.rel_mse(ia, ib)
    sc.close()
    row = {"config": name, "triangles": int(len(flat.tri_indices)), "instances": int(len(flat.instances)), "lights": int(len(flat.lights)), "host_build_s": round(t_build, 2),
           "spp_step": spp, "gpu_ms": round(ms, 2), "gpu_samples_per_s": st.camera_rays / ms * 1e3, "gpu_mrays_per_s": rays / ms / 1e3, "rays_per_sample": rays / max(st.camera_rays, 1),
           "cpu_samples_per_s": cpu_sps, "cpu_threads": nth, "speedup_vs_cpu_port": st.camera_rays / ms * 1e3 / cpu_sps, "relmse_crop_4spp": rel, "iterations": int(st.iterations)}
    out.append(row)
    print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/config_sweep.json", "w"), indent=1)
