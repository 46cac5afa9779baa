import importlib, sys, time
sys.path.insert(0, '.')
import numpy as np
P = importlib.import_module("pbrt-rust_b200")
from oracle import oracle as O
import torch
S = P.scenes
setup = S.displaced_sphere_scene()
flat = setup.flat
sc = P.Scene(flat)
integ = setup.make_integrator()
for name, rays in (("cam", S.rays_camera(integ)), ("diff", S.rays_diffuse(flat, 4_000_000))):
    n = len(rays)
    t0 = time.time(); want, cnt = O.intersect(flat, rays[:200000]); cpu_s = time.time() - t0
    got = sc.intersect(rays)
    print(name, "parity", got[:200000].tobytes() == want.tobytes(), "hit frac", (got['prim'] != 0xffffffff).mean(), "cpu Mrays/s", 0.2/cpu_s, "nodes/ray", cnt[0]/cnt[2], "tris/ray", cnt[1]/cnt[2])
    dr = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    dh = torch.empty((n, 4), dtype=torch.int32, device='cuda')
    for _ in range(3): sc.intersect_dev(dr.data_ptr(), n, dh.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): sc.intersect_dev(dr.data_ptr(), n, dh.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, n, "rays", ms, "ms", n / ms / 1e3, "Mrays/s")
