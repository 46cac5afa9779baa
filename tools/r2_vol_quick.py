"""volpath throughput on the three scenes of tools/r2_f4_sweep.py (GPU only, one 16-spp render after a warm-up): A/B of k_vol_mega changes."""
import importlib, sys
sys.path.insert(0, '.')
import torch
P = importlib.import_module("pbrt-rust_b200")
S, H = P.scenes, P.host


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
    sc.close()
    print(f"{name}: {st.camera_rays / ms / 1e3:.1f} M samples/s, {st.intersection_tests / ms / 1e3:.0f} Mrays/s, {ms:.1f} ms", flush=True)


fog = S.fog_box_scene(xres=1024, yres=1024, spp=64)
measure("volpath fog box d5", fog.flat, fog.make_integrator(spp_=64), 16)
measure("volpath fog box d40", fog.flat, fog.make_integrator(spp_=64, maxdepth_=40), 16)
s3 = S.displaced_sphere_scene()
base = s3.make_integrator(spp_=64)
measure("volpath S3", s3.flat, H.VolPathIntegrator(base.camera, base.film, base.sampler, maxdepth=5, lightsamplestrategy="power"), 16)
t1 = S.textured_scene(xres=1920, yres=1080, spp=64)
measure("volpath T1", t1.flat, t1.make_integrator(integrator="volpath", spp_=64), 8)
if len(sys.argv) > 1 and sys.argv[1] == "zt":
    c = S.cornell_scene(xres=512, yres=512, spp=16, sampler="02sequence")
    measure("path 02sequence cornell 512", c.flat, c.make_integrator(), 8)
    t = S.textured_scene(xres=640, yres=480, spp=8, sampler="02sequence")
    measure("path 02sequence T1 640x480", t.flat, t.make_integrator(), 4)
