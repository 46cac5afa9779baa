"""A/B timing of the batch closest-hit kernel variants on incoherent surface rays (S3)."""
import importlib, sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np
import torch
P = importlib.import_module("pbrt-rust_b200")
S = P.scenes
setup = S.displaced_sphere_scene()
flat = setup.flat
sc = P.Scene(flat)
lib = P.load_library()
lib.pbrt_b200_debug_tune.argtypes = [C.c_int, C.c_int]
n = 4_000_000
batches = {"surf": S.rays_surface(flat, n, tri_hi=len(flat.tri_indices) - 4), "cam": S.rays_camera(setup.make_integrator())}
def run(rays, reps=10):
    m = len(rays)
    dr = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    dh = torch.empty((m, 4), dtype=torch.int32, device='cuda')
    for _ in range(3): sc.intersect_dev(dr.data_ptr(), m, dh.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): sc.intersect_dev(dr.data_ptr(), m, dh.data_ptr())
    e1.record(); torch.cuda.synchronize()
    return m / (e0.elapsed_time(e1) / reps) / 1e3, dh.cpu().numpy()
configs = [("ifif 1rpt", {0: 2}), ("whilewhile 1rpt", {0: 1})]
for rb in (0, 8, 16, 20, 24, 28):
    for ch in (32, 96, 256):
        configs.append((f"persistent refill<{rb} chunk {ch}", {0: 0, 1: rb, 2: ch}))
ref = {}
for name, tune in configs:
    for k, v in {0: 0, 1: 20, 2: 96, 3: 0, **tune}.items(): lib.pbrt_b200_debug_tune(k, v)
    out = []
    for bn, rays in batches.items():
        mr, hits = run(rays)
        if bn not in ref: ref[bn] = hits
        out.append(f"{bn} {mr:8.1f} Mrays/s same={np.array_equal(ref[bn], hits)}")
    print(f"{name:36s}", " | ".join(out), flush=True)
