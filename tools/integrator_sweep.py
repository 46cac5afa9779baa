"""SURVEY.md §8 f2 / f4 measurements on one B200 (run under gpurun; writes gpurun_out/integrator_sweep.json):
* DirectLightingIntegrator ("one", "all") and WhittedIntegrator next to the PathIntegrator on C1 / C2 / C3 scenes:
  device-resident samples/s and Mrays/s, the CPU oracle's samples/s on a bounded sample, image relMSE on a crop;
* the scene-file front end on S3: write .pbrt + .ply, parse + flatten (+ BVH build), tables identical or not."""
import importlib, json, os, sys, tempfile, time
sys.path.insert(0, '.')
import numpy as np
import torch
P = importlib.import_module("pbrt-rust_b200")
SF = importlib.import_module("pbrt-rust_b200.scenefile")
from oracle import oracle as O
S, H = P.scenes, P.host
nth = os.cpu_count() or 1
out = {"integrators": [], "frontend": {}}
SCENES = [("C1 spheres 400x400", S.spheres_scene, 64), ("C2 cornell 1024x1024", S.cornell_scene, 16), ("C3 S3 1M-tri 1920x1080", S.displaced_sphere_scene, 8)]
for sname, make, spp in SCENES:
    setup = make()
    flat = setup.flat
    base = setup.make_integrator(spp_=spp * 3)
    variants = [("path d5", H.PathIntegrator(base.camera, base.film, base.sampler, maxdepth=5, lightsamplestrategy="power")),
                ("directlighting one d5", H.DirectLightingIntegrator(base.camera, base.film, base.sampler, maxdepth=5, strategy="one")),
                ("directlighting all d5", H.DirectLightingIntegrator(base.camera, base.film, base.sampler, maxdepth=5, strategy="all")),
                ("whitted d5", H.WhittedIntegrator(base.camera, base.film, base.sampler, maxdepth=5))]
    sc = P.Scene(flat)
    film = base.film
    film_t = torch.zeros((film.width * film.height, 4), dtype=torch.float32, device="cuda")
    for vname, integ in variants:
        sc.render(integ, sample_range=(0, spp), device_ptr=film_t.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        film_t.zero_(); e0.record()
        _, st = sc.render(integ, sample_range=(spp, 2 * spp), device_ptr=film_t.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rays = st.intersection_tests + st.shadow_tests
        nt = integ.n_tiles()
        win = (0, nt) if film.width * film.height <= 1100 * 1100 else (nt // 3, nt // 3 + max(nt // 12, 1))
        t0 = time.time(); _, ost = O.render(flat, integ, nthreads=nth, tile_range=win, sample_range=(0, 2)); t_cpu = time.time() - t0
        crop = (nt // 2, nt // 2 + min(nt // 2, 96))
        a, _ = sc.render(integ, tile_range=crop, sample_range=(0, 4))
        b, _ = O.render(flat, integ, nthreads=nth, tile_range=crop, sample_range=(0, 4))
        m = b[:, 3] > 0
        rel = O.rel_mse(sc.film_resolve(a[m], film.scale), O.film_resolve(b[m], film.scale))
        row = {"scene": sname, "integrator": vname, "spp_step": spp, "gpu_ms": round(ms, 2), "gpu_samples_per_s": st.camera_rays / ms * 1e3, "gpu_mrays_per_s": rays / ms / 1e3,
               "rays_per_sample": rays / max(st.camera_rays, 1), "cpu_samples_per_s": ost["camera_rays"] / t_cpu, "cpu_threads": nth,
               "speedup_vs_cpu_port": st.camera_rays / ms * 1e3 / (ost["camera_rays"] / t_cpu), "relmse_crop_4spp": rel, "iterations": int(st.iterations)}
        out["integrators"].append(row)
        print(json.dumps(row), flush=True)
    sc.close()
    if "S3" in sname:
        d = tempfile.mkdtemp()
        integ = setup.make_integrator()
        t0 = time.time(); files = SF.write_pbrt(os.path.join(d, "s3.pbrt"), flat, integ); t1 = time.time()
        job = P.pbrt_parse(files[0]).jobs[0]; t2 = time.time()
        same = all((getattr(job.flat, k) is None and getattr(flat, k) is None) or getattr(job.flat, k).tobytes() == getattr(flat, k).tobytes()
                   for k in ("nodes", "prims", "vertex_p", "vertex_n", "vertex_uv", "tri_indices", "materials", "lights"))
        t3 = time.time(); img, st = job.render(device=0, sample_range=(0, 4)); t4 = time.time()
        out["frontend"] = {"scene": sname, "files": len(files), "bytes": int(sum(os.path.getsize(f) for f in files)), "write_s": round(t1 - t0, 3),
                           "parse_flatten_bvh_s": round(t2 - t1, 3), "tables_identical": bool(same), "render_4spp_e2e_s": round(t4 - t3, 3),
                           "triangles": int(len(flat.tri_indices))}
        print(json.dumps(out["frontend"]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/integrator_sweep.json", "w"), indent=1)
