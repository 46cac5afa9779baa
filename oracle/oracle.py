"""TEST INFRASTRUCTURE -- ctypes binding of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
_lib = None


def build(force=False):
    srcs = list(HERE.glob("*.hpp")) + [HERE / "capi.cpp", HERE.parent / "include" / "pbrt_b200.h"]
    if force or not LIB.exists() or LIB.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.orc_next_float_up.restype = C.c_float
        _lib.orc_next_float_up.argtypes = [C.c_float]
        _lib.orc_next_float_down.restype = C.c_float
        _lib.orc_next_float_down.argtypes = [C.c_float]
        _lib.orc_gamma.restype = C.c_float
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bvh_build(prim_bounds, max_prims=4, split_method="sah"):
    import importlib
    H = importlib.import_module("pbrt-rust_b200.host")
    pb = np.ascontiguousarray(prim_bounds, np.float32).reshape(-1, 6)
    n = len(pb)
    nodes = np.zeros(max(2 * n - 1, 1), H.NODE_DTYPE)
    ordered = np.zeros(max(n, 1), np.uint32)
    nn = C.c_uint64(0)
    lib().orc_bvh_build(ptr(pb), C.c_uint64(n), int(max_prims), H.SPLIT[split_method], ptr(nodes), ptr(ordered), C.byref(nn))
    return nodes[: nn.value].copy(), ordered[:n].copy()


def intersect(flat, rays, nthreads=0, per_ray=False):
    import importlib
    H = importlib.import_module("pbrt-rust_b200.host")
    rays = np.ascontiguousarray(rays, H.RAY_DTYPE)
    hits = np.zeros(len(rays), H.HIT_DTYPE)
    counters = np.zeros(3, np.uint64)
    prc = np.zeros((len(rays), 2), np.uint32) if per_ray else None
    d = flat.desc()
    lib().orc_intersect(C.byref(d), ptr(rays), C.c_uint64(len(rays)), ptr(hits), int(nthreads), ptr(counters), ptr(prc))
    return (hits, counters, prc) if per_ray else (hits, counters)


def intersect_p(flat, rays, nthreads=0):
    import importlib
    H = importlib.import_module("pbrt-rust_b200.host")
    rays = np.ascontiguousarray(rays, H.RAY_DTYPE)
    occ = np.zeros(len(rays), np.uint8)
    counters = np.zeros(3, np.uint64)
    d = flat.desc()
    lib().orc_intersect_p(C.byref(d), ptr(rays), C.c_uint64(len(rays)), ptr(occ), int(nthreads), ptr(counters))
    return occ.astype(bool), counters
