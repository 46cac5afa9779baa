"""The C-ABI shared library loads, exports every symbol include/pbrt_b200.h declares, and (on a box without a
GPU) refuses compute calls loudly instead of falling back to a CPU path."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "pbrt_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(pbrt_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("pbrt_b200_scene_create", "pbrt_b200_intersect", "pbrt_b200_intersect_p", "pbrt_b200_render", "pbrt_b200_film_resolve", "pbrt_b200_bvh_build"):
        assert s in syms


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"libpbrt_b200.so lacks {missing}"
    assert sorted(pkg.host.EXPORTS) == declared_symbols(), "host.EXPORTS and the header disagree"
    # v2: object instancing tables in pbrt_b200_scene_desc; v3: pbrt_b200_light.n_samples; v4: media tables, integrator.camera_medium
    assert lib.pbrt_b200_abi_version() == 5


def test_struct_layouts_match_the_header(pkg):
    H = pkg.host
    # sizes stated in include/pbrt_b200.h
    assert H.NODE_DTYPE.itemsize == 32 and H.PRIM_DTYPE.itemsize == 24 and H.MATERIAL_DTYPE.itemsize == 48
    assert H.LIGHT_DTYPE.itemsize == 136 and H.RAY_DTYPE.itemsize == 32 and H.HIT_DTYPE.itemsize == 16 and H.SPHERE_DTYPE.itemsize == 144
    lib = pkg.load_library()
    if hasattr(lib, "pbrt_b200_struct_size"):
        for name, py in (("scene_desc", C.sizeof(H.SceneDesc)), ("render_desc", C.sizeof(H.RenderDesc)), ("render_stats", C.sizeof(H.RenderStats))):
            assert lib.pbrt_b200_struct_size(name.encode()) == py, name


def test_no_cpu_fallback_without_a_device(pkg):
    lib = pkg.load_library()
    if lib.pbrt_b200_device_count() > 0:
        pytest.skip("a CUDA device is visible; the refusal path is for GPU-less hosts")
    flat = pkg.scenes.small_mixed_scene().flat
    with pytest.raises(pkg.B200Error) as e:
        pkg.Scene(flat)
    assert "code 2" in str(e.value)  # PBRT_B200_ERR_NO_DEVICE
    assert lib.pbrt_b200_last_error()


def test_invalid_arguments_are_reported(pkg):
    lib = pkg.load_library()
    out = C.c_void_p()
    assert lib.pbrt_b200_scene_create(None, 0, C.byref(out)) != 0
    assert b"" != lib.pbrt_b200_last_error()
    assert lib.pbrt_b200_film_resolve(None, 4, C.c_float(1.0), None) != 0


def test_malformed_scene_tables_are_rejected_on_the_host(pkg):
    """pbrt_b200_scene_create validates the tables before anything reaches the device (so this runs without a GPU): null tables with a
    non-zero count, and a LinearBVHNode array that is not a tree -- interior nodes sharing a child pass every per-node check, but the
    device-side layout build sizes its arrays for a tree."""
    lib = pkg.load_library()
    H = pkg.host

    def rc_of(flat, mutate):
        d = flat.desc()
        keep = mutate(d, flat)  # keep any replacement arrays alive until the call returns
        out = C.c_void_p()
        rc = lib.pbrt_b200_scene_create(C.byref(d), 0, C.byref(out))
        if rc == 0:
            lib.pbrt_b200_scene_destroy(out)
        del keep
        return rc, lib.pbrt_b200_last_error().decode()

    flat = pkg.scenes.small_mixed_scene().flat
    for field in ("tri_indices", "vertex_p", "materials", "lights"):
        def null_it(d, f, field=field):
            setattr(d, field, None)
        rc, msg = rc_of(flat, null_it)
        assert rc == 1 and field in msg, (field, rc, msg)  # PBRT_B200_ERR_INVALID

    def dag(d, f):
        # a chain of interior nodes that all take the LAST node (a leaf) as their second child: every node has c1 > c0, in range, axis ok
        nodes = f.nodes.copy()
        nn = len(nodes)
        assert nn >= 8
        last = nn - 1
        assert nodes[last]["n_prims"] != 0
        for i in range(nn - 2):
            nodes[i]["n_prims"] = 0
            nodes[i]["offset"] = last
            nodes[i]["axis"] = 0
        d.nodes = nodes.ctypes.data
        return nodes

    rc, msg = rc_of(flat, dag)
    assert rc == 1 and ("not a tree" in msg or "malformed" in msg or "deeper" in msg), (rc, msg)


def test_film_resolve_matches_oracle(pkg, oracle):
    """Film::write_image arithmetic (film.rs:217-264) is host code in the product library: compare with the oracle."""
    rng = np.random.RandomState(0)
    rgbw = rng.uniform(0, 4, size=(1000, 4)).astype(np.float32)
    rgbw[::7, 3] = 0.0          # zero-weight pixels are not divided
    rgbw[::11, :3] *= -1.0      # negative values clamp at 0 after the division
    lib = pkg.load_library()
    out = np.zeros((1000, 3), np.float32)
    assert lib.pbrt_b200_film_resolve(rgbw.ctypes.data_as(C.c_void_p), 1000, C.c_float(0.5), out.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(out, oracle.film_resolve(rgbw, 0.5))
