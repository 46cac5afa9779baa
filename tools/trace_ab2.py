"""Potential of ray reordering: bounce-1 rays in pixel order vs. octant-grouped vs. fully sorted."""
import importlib, sys, time, ctypes as C, math
sys.path.insert(0, '.')
import numpy as np
import torch
P = importlib.import_module("pbrt-rust_b200")
S = P.scenes
H = P.host
setup = S.displaced_sphere_scene()
flat = setup.flat
sc = P.Scene(flat)
lib = P.load_library()
lib.pbrt_b200_debug_tune.argtypes = [C.c_int, C.c_int]
cam = S.rays_camera(setup.make_integrator())
# tile-order the camera rays like the wavefront does (16x16 tiles, x fastest)
W, Hh = 1920, 1080
ids = np.arange(W * Hh).reshape(Hh, W)
pad = np.full(((Hh + 15) // 16 * 16, W), -1); pad[:Hh] = ids
tiles = pad.reshape(-1, 16, W // 16, 16).transpose(0, 2, 1, 3).reshape(-1)
cam = cam[tiles[tiles >= 0]]
hits = sc.intersect(cam)
ok = hits["prim"] != H.NO_HIT
cam, hits = cam[ok], hits[ok]
p = cam["o"] + cam["d"] * hits["t"][:, None]
# normals from the hit triangle
slot_of = np.zeros(len(flat.prims), np.int64); slot_of[flat.prims["creation_index"]] = np.arange(len(flat.prims))
tri = flat.prims["shape_index"][slot_of[hits["prim"]]]
idx = flat.tri_indices[tri]
p0, p1, p2 = flat.vertex_p[idx[:, 0]], flat.vertex_p[idx[:, 1]], flat.vertex_p[idx[:, 2]]
n = np.cross(p1 - p0, p2 - p0); n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
n[(n * cam["d"]).sum(1) > 0] *= -1
u = S._hash_floats(2 * len(p), 5).reshape(2, -1)
z = 1 - 2 * u[0]; r = np.sqrt(np.maximum(0, 1 - z * z)); phi = 2 * math.pi * u[1]
d = np.stack([r * np.cos(phi), r * np.sin(phi), z], 1).astype(np.float32)
d[(d * n).sum(1) < 0] *= -1
b1 = H.make_rays((p + n * 1e-4).astype(np.float32), d)
print("bounce-1 rays", len(b1))
octant = ((d[:, 0] < 0).astype(np.int64) | ((d[:, 1] < 0).astype(np.int64) << 1) | ((d[:, 2] < 0).astype(np.int64) << 2))
def block_sort(keys, block):
    nblk = (len(keys) + block - 1) // block
    blk = np.arange(len(keys)) // block
    return np.lexsort((np.arange(len(keys)), keys, blk))
lo, hi = p.min(0), p.max(0)
cell = np.minimum(((p - lo) / (hi - lo + 1e-9) * 64).astype(np.int64), 63)
def morton(c):
    m = np.zeros(len(c), np.int64)
    for b in range(6):
        for k in range(3): m |= ((c[:, k] >> b) & 1) << (3 * b + k)
    return m
mort = morton(cell)
orders = {"pixel order": np.arange(len(b1)), "octant within 1024": block_sort(octant, 1024), "octant within 8192": block_sort(octant, 8192),
          "octant within 65536": block_sort(octant, 65536), "global (octant, morton18)": np.lexsort((mort, octant)),
          "global (morton9, octant)": np.lexsort((octant, mort >> 9)), "random shuffle": np.random.default_rng(0).permutation(len(b1))}
def run(rays, reps=10):
    m = len(rays)
    dr = torch.from_numpy(np.ascontiguousarray(rays).view(np.float32).reshape(-1, 8)).cuda()
    dh = torch.empty((m, 4), dtype=torch.int32, device='cuda')
    for _ in range(3): sc.intersect_dev(dr.data_ptr(), m, dh.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): sc.intersect_dev(dr.data_ptr(), m, dh.data_ptr())
    e1.record(); torch.cuda.synchronize()
    return m / (e0.elapsed_time(e1) / reps) / 1e3
for vname, tune in (("ifif", {0: 2}), ("persistent r24 c32", {0: 0, 1: 24, 2: 32})):
    for k, v in {0: 0, 1: 20, 2: 96, 3: 0, **tune}.items(): lib.pbrt_b200_debug_tune(k, v)
    for oname, order in orders.items():
        print(f"{vname:20s} {oname:28s} {run(b1[order]):8.1f} Mrays/s", flush=True)
