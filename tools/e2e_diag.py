"""Per-phase host timings of the end-to-end call sequence bench.py's `e2e` leg times (scene_create + render + destroy),
printed by the library itself when PBRT_B200_PROFILE is set.  Run under gpurun."""
import importlib, os, sys, time
os.environ["PBRT_B200_PROFILE"] = "1"
sys.path.insert(0, '.')
import numpy as np
P = importlib.import_module("pbrt-rust_b200")
setup = P.scenes.displaced_sphere_scene()
integ = setup.make_integrator(spp_=16 * 8)
film = integ.film
host_film = np.zeros((film.width * film.height, 4), np.float32)
if os.environ.get("PINNED"):
    import copy, torch
    keep = []
    def pinned(a):
        if a is None or a.nbytes == 0:
            return a
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory(); keep.append(t)
        v = t.numpy().view(a.dtype).reshape(a.shape); v[...] = a
        return v
    flat = copy.copy(setup.flat)
    for name in ("nodes", "prims", "vertex_p", "vertex_n", "vertex_s", "vertex_uv", "tri_indices", "spheres", "materials", "lights"):
        setattr(flat, name, pinned(getattr(flat, name)))
    setup.flat = flat
    host_film = pinned(host_film)
FLAGS = P.host.RENDER_OVERWRITE if os.environ.get("PINNED") else 0
for k in range(4):
    t0 = time.perf_counter()
    sc = P.Scene(setup.flat)
    t1 = time.perf_counter()
    host_film[:] = 0
    t2 = time.perf_counter()
    _, st = sc.render(integ, rgbw=host_film, sample_range=(k * 16, (k + 1) * 16), flags=FLAGS)
    t3 = time.perf_counter()
    sc.close()
    t4 = time.perf_counter()
    print(f"step {k}: create {1e3 * (t1 - t0):.1f} ms, zero film {1e3 * (t2 - t1):.1f}, render {1e3 * (t3 - t2):.1f} (device {st.device_ms:.1f}), destroy {1e3 * (t4 - t3):.1f}, "
          f"total {1e3 * (t4 - t0):.1f} ms -> {st.camera_rays / (t4 - t0) / 1e6:.1f} Msamples/s", file=sys.stderr, flush=True)
