python - <<'PY'
import importlib, sys, time, os
sys.path.insert(0,'.')
import numpy as np
P = importlib.import_module("pbrt-rust_b200")
ref = {}
for name, setup, kw in (("C1 spheres 400x400 16spp", P.scenes.spheres_scene(), dict(spp_=16)), ("S3 1080p 4spp", P.scenes.displaced_sphere_scene(), dict(spp_=4)), ("S3 4K 2spp", P.scenes.displaced_sphere_scene(), dict(spp_=2, res=(3840, 2160)))):
    integ = setup.make_integrator(sampler_="02sequence", **kw)
    sc = P.Scene(setup.flat)
    for lanes in (1, 2, 4, 8, 16, 32, 0):
        if lanes: os.environ["PBRT_B200_ZT_LANES"] = str(lanes)
        else: os.environ.pop("PBRT_B200_ZT_LANES", None)
        img, st = integ.render(sc)
        if lanes == 1: ref[name] = img
        print(name, "lanes", lanes or "auto", "device %.1f ms"%st.device_ms, "%.2f M samples/s"%(st.camera_rays/st.device_ms/1e3), "same image", bool(np.allclose(img, ref[name], rtol=2e-5, atol=2e-5)), flush=True)
    sc.close()
PY
