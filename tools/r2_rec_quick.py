"""whitted / directlighting throughput on the textured scene at 1920x1080 (GPU only, one 8-spp render after a warm-up): A/B of k_rec_shade<.., TEX>."""
import importlib, sys
sys.path.insert(0, '.')
import torch
P = importlib.import_module("pbrt-rust_b200")
t1 = P.scenes.textured_scene(xres=1920, yres=1080, spp=64)
sc = P.Scene(t1.flat)
for name in ("whitted", "directlighting:all", "directlighting:one"):
    integ = t1.make_integrator(integrator=name, spp_=64)
    film = integ.film
    film_t = torch.zeros((film.width * film.height, 4), dtype=torch.float32, device="cuda")
    sc.render(integ, sample_range=(0, 8), device_ptr=film_t.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    film_t.zero_(); e0.record()
    _, st = sc.render(integ, sample_range=(8, 16), device_ptr=film_t.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name} T1: {st.camera_rays / ms / 1e3:.1f} M samples/s, {ms:.1f} ms, shade {st.shade_ms:.1f} ms, launches {st.kernel_launches}", flush=True)
sc.close()
