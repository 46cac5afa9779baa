"""TEST INFRASTRUCTURE -- ctypes binding of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes as C
import importlib
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
_lib = None


def _H():
    return importlib.import_module("pbrt-rust_b200.host")


def build(force=False):
    srcs = list(HERE.glob("*.hpp")) + [HERE / "capi.cpp", HERE.parent / "include" / "pbrt_b200.h"]
    if force or not LIB.exists() or LIB.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(HERE)] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        for name in ("orc_next_float_up", "orc_next_float_down"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_float]
        L.orc_gamma.restype = C.c_float
        L.orc_sobol_sample_float.restype = C.c_float
        L.orc_sobol_sample_float.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32]
        L.orc_sobol_interval_to_index.restype = C.c_uint64
        L.orc_sobol_interval_to_index.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int]
        L.orc_radical_inverse.restype = C.c_float
        L.orc_radical_inverse.argtypes = [C.c_int, C.c_uint64]
        L.orc_scrambled_radical_inverse.restype = C.c_float
        L.orc_scrambled_radical_inverse.argtypes = [C.c_int, C.c_uint64]
        L.orc_inverse_radical_inverse.restype = C.c_uint64
        L.orc_inverse_radical_inverse.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_rng_u32.restype = C.c_uint32
        L.orc_rng_u32.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.orc_film_resolve.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p]
        L.orc_film_resolve.restype = None
        L.orc_distribution1d.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.orc_find_interval.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.orc_triangle_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.orc_sphere_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_efloat_op.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_sampler_stream.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bvh_build(prim_bounds, max_prims=4, split_method="sah"):
    H = _H()
    pb = np.ascontiguousarray(prim_bounds, np.float32).reshape(-1, 6)
    n = len(pb)
    nodes = np.zeros(max(2 * n - 1, 1), H.NODE_DTYPE)
    ordered = np.zeros(max(n, 1), np.uint32)
    nn = C.c_uint64(0)
    lib().orc_bvh_build(ptr(pb), C.c_uint64(n), int(max_prims), H.SPLIT[split_method], ptr(nodes), ptr(ordered), C.byref(nn))
    return nodes[: nn.value].copy(), ordered[:n].copy()


def intersect(flat, rays, nthreads=0, per_ray=False):
    H = _H()
    rays = np.ascontiguousarray(rays, H.RAY_DTYPE)
    hits = np.zeros(len(rays), H.HIT_DTYPE)
    counters = np.zeros(3, np.uint64)
    prc = np.zeros((len(rays), 2), np.uint32) if per_ray else None
    d = flat.desc()
    lib().orc_intersect(C.byref(d), ptr(rays), C.c_uint64(len(rays)), ptr(hits), int(nthreads), ptr(counters), ptr(prc))
    return (hits, counters, prc) if per_ray else (hits, counters)


def intersect_p(flat, rays, nthreads=0):
    H = _H()
    rays = np.ascontiguousarray(rays, H.RAY_DTYPE)
    occ = np.zeros(len(rays), np.uint8)
    counters = np.zeros(3, np.uint64)
    d = flat.desc()
    lib().orc_intersect_p(C.byref(d), ptr(rays), C.c_uint64(len(rays)), ptr(occ), int(nthreads), ptr(counters))
    return occ.astype(bool), counters


STAT_NAMES = ["camera_rays", "intersection_tests", "shadow_tests", "zero_radiance", "direct_den", "closest_nodes", "closest_prims", "closest_rays",
              "any_nodes", "any_prims", "any_rays", "reserved"]


def render(flat, integrator, nthreads=0, tile_range=None, sample_range=None, rgbw=None, tile_interleave=None, tile_order=0):
    """SamplerIntegrator::render on the CPU oracle -> ({r,g,b,w} sums [npix,4], stats dict)."""
    film = integrator.film
    if rgbw is None:
        rgbw = np.zeros((film.height * film.width, 4), np.float32)
    sd = flat.desc()
    rd = integrator.desc(tile_range, sample_range, tile_interleave=tile_interleave, tile_order=tile_order)
    stats = np.zeros(12, np.uint64)
    lib().orc_render(C.byref(sd), C.byref(rd), ptr(rgbw), int(nthreads), ptr(stats))
    return rgbw, dict(zip(STAT_NAMES, (int(v) for v in stats)))


def film_resolve(rgbw, scale=1.0):
    rgbw = np.ascontiguousarray(rgbw, np.float32).reshape(-1, 4)
    out = np.zeros((len(rgbw), 3), np.float32)
    lib().orc_film_resolve(ptr(rgbw), len(rgbw), scale, ptr(out))
    return out


def render_image(flat, integrator, **kw):
    rgbw, stats = render(flat, integrator, **kw)
    film = integrator.film
    return film_resolve(rgbw, film.scale).reshape(film.height, film.width, 3), stats


def sampler_stream(integrator, seed, px, py, nsamples, n1d2d, prefix_pixels=()):
    sd = integrator.sampler.desc(integrator.film)
    pre = np.ascontiguousarray(np.array(prefix_pixels, np.int32).reshape(-1, 2))
    out = np.zeros(nsamples * (5 + 3 * n1d2d), np.float32)
    k = lib().orc_sampler_stream(C.byref(sd), int(seed), ptr(pre) if len(pre) else None, len(pre), int(px), int(py), int(nsamples), int(n1d2d), ptr(out))
    return out[:k].reshape(-1, 5 + 3 * n1d2d)


def bsdf_eval(material_row, wo, wi, u, flags=31):
    m = np.array([material_row], dtype=_H().MATERIAL_DTYPE)
    wo, wi, u = (np.ascontiguousarray(v, np.float32) for v in (wo, wi, u))
    out = np.zeros(13, np.float32)
    lib().orc_bsdf_eval(ptr(m), ptr(wo), ptr(wi), ptr(u), int(flags), ptr(out))
    return out


def bsdf_eval_batch(material_row, wo, wi, u, flags=31):
    """n evaluations for one material / wo: wi [n,3], u [n,2] -> [n,13] rows laid out like bsdf_eval's.
    `material_row`: a MATERIAL_DTYPE row, or a host.TexturedMaterial whose parameters are all constants (uber, substrate)."""
    H = _H()
    ext = None
    if isinstance(material_row, H.TexturedMaterial):
        T = importlib.import_module("pbrt-rust_b200.textures")
        tm = material_row
        ext = np.zeros(1, T.MATERIAL_EXT_DTYPE)
        for k, v in enumerate(tm.spectra):
            ext[0]["s_const"][k] = np.full(3, v, np.float32) if np.ndim(v) == 0 else np.asarray(v, np.float32)
        for k, v in enumerate(tm.floats):
            ext[0]["f_const"][k] = np.float32(v)
        assert tm.bump is None
        material_row = tm.row
    m = np.array([material_row], dtype=H.MATERIAL_DTYPE)
    wo = np.ascontiguousarray(wo, np.float32)
    wi = np.ascontiguousarray(wi, np.float32).reshape(-1, 3)
    u = np.ascontiguousarray(u, np.float32).reshape(-1, 2)
    assert len(wi) == len(u)
    out = np.zeros((len(wi), 13), np.float32)
    if ext is None:
        lib().orc_bsdf_eval_batch(ptr(m), ptr(wo), ptr(wi), ptr(u), int(flags), C.c_uint64(len(wi)), ptr(out))
    else:
        lib().orc_bsdf_eval_batch_ext(ptr(m), ptr(ext), ptr(wo), ptr(wi), ptr(u), int(flags), C.c_uint64(len(wi)), ptr(out))
    return out


def generate_ray_differential(camera, pfilm, time=0.0, plens=(0.0, 0.0), spp=0):
    """-> [6,3]: o, d, rx_origin, rx_direction, ry_origin, ry_direction of the camera ray through `pfilm`."""
    cd = camera.desc()
    cs = np.array([pfilm[0], pfilm[1], time, plens[0], plens[1]], np.float32)
    out = np.zeros(18, np.float32)
    lib().orc_generate_ray_differential(C.byref(cd), ptr(cs), C.c_uint32(int(spp)), ptr(out))
    return out.reshape(6, 3)


def light_sample_batch(flat, light, ref_p, ref_n, u):
    """Light::sample_li from (ref_p, ref_n) for u [n,2] -> [n,8] rows {Li.rgb, wi.xyz, pdf, pdf_li(wi)}."""
    u = np.ascontiguousarray(u, np.float32).reshape(-1, 2)
    out = np.zeros((len(u), 8), np.float32)
    d = flat.desc()
    rc = lib().orc_light_sample_batch(C.byref(d), int(light), ptr(np.ascontiguousarray(ref_p, np.float32)), ptr(np.ascontiguousarray(ref_n, np.float32)), ptr(u),
                                      C.c_uint64(len(u)), ptr(out))
    if rc:
        raise ValueError("light index out of range")
    return out


def texture_eval(flat, texref, p, uv=None, dpdx=None, dpdy=None, duv=None):
    """Texture::evaluate of the flattened program `texref` = (first, count) at points p [n,3] (uv [n,2], dpdx / dpdy [n,3],
    duv [n,4] = dudx dvdx dudy dvdy; zeros when omitted) -> [n,3]."""
    p = np.ascontiguousarray(p, np.float32).reshape(-1, 3)
    n = len(p)
    q = np.zeros((n, 16), np.float32)
    q[:, 0:3] = p
    for lo, hi, v in ((3, 5, uv), (5, 8, dpdx), (8, 11, dpdy), (11, 15, duv)):
        if v is not None:
            q[:, lo:hi] = np.asarray(v, np.float32).reshape(-1, hi - lo)
    out = np.zeros((n, 3), np.float32)
    d = flat.desc()
    lib().orc_texture_eval(C.byref(d), C.c_uint32(int(texref[0])), C.c_uint32(int(texref[1])), C.c_uint64(n), ptr(q), ptr(out))
    return out


def rel_mse(img, ref):
    """relMSE of SURVEY.md s8(d): mean over pixels/channels of (a-b)^2 / (b^2 + 1e-2)."""
    a, b = np.asarray(img, np.float64), np.asarray(ref, np.float64)
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def spatial_lookup(flat, points):
    """SpatialLightDistribution::lookup at `points` -> (voxel [n,3] int32, func [n,n_lights] f32, nvoxels [3])."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    d = flat.desc()
    nl = int(d.n_lights)
    voxel = np.zeros((len(pts), 3), np.int32)
    func = np.zeros((len(pts), max(nl, 1)), np.float32)
    nv = np.zeros(3, np.int32)
    lib().orc_spatial_lookup(C.byref(d), ptr(pts), C.c_uint64(len(pts)), ptr(voxel), ptr(func), ptr(nv))
    return voxel, func[:, :nl], nv
