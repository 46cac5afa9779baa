import importlib, sys
sys.path.insert(0, '.')
import torch
P = importlib.import_module("pbrt-rust_b200")
S, H = P.scenes, P.host
s3 = S.displaced_sphere_scene()
base = s3.make_integrator(spp_=64)
sc = P.Scene(s3.flat)
for name, integ in (("whitted", H.WhittedIntegrator(base.camera, base.film, base.sampler, maxdepth=5)),
                    ("directlighting all", H.DirectLightingIntegrator(base.camera, base.film, base.sampler, maxdepth=5, strategy="all")),
                    ("directlighting one", H.DirectLightingIntegrator(base.camera, base.film, base.sampler, maxdepth=5, strategy="one"))):
    film = integ.film
    film_t = torch.zeros((film.width * film.height, 4), dtype=torch.float32, device="cuda")
    sc.render(integ, sample_range=(0, 16), device_ptr=film_t.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    film_t.zero_(); e0.record()
    _, st = sc.render(integ, sample_range=(16, 32), device_ptr=film_t.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name} S3: {st.camera_rays / ms / 1e3:.1f} M samples/s, {ms:.1f} ms, shade {st.shade_ms:.1f}", flush=True)
