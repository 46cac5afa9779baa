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


def test_struct_layouts_match_what_a_c_compiler_sees(pkg, tmp_path):
    """include/pbrt_b200.h compiled as C by gcc: sizeof / offsetof of every struct that crosses the boundary equal the numpy dtypes and
    ctypes Structures the Python host fills (a field added on one side only shifts everything behind it silently)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    H = pkg.host
    T = __import__("importlib").import_module("pbrt-rust_b200.textures")
    structs = {"pbrt_b200_bvh_node": H.NODE_DTYPE, "pbrt_b200_prim": H.PRIM_DTYPE, "pbrt_b200_sphere": H.SPHERE_DTYPE, "pbrt_b200_material": H.MATERIAL_DTYPE,
               "pbrt_b200_light": H.LIGHT_DTYPE, "pbrt_b200_ray": H.RAY_DTYPE, "pbrt_b200_hit": H.HIT_DTYPE, "pbrt_b200_object": H.OBJECT_DTYPE,
               "pbrt_b200_instance": H.INSTANCE_DTYPE, "pbrt_b200_medium": H.MEDIUM_DTYPE, "pbrt_b200_medium_interface": H.MEDIUM_INTERFACE_DTYPE,
               "pbrt_b200_texnode": T.TEXNODE_DTYPE, "pbrt_b200_mipmap": T.MIPMAP_DTYPE, "pbrt_b200_texref": T.TEXREF_DTYPE,
               "pbrt_b200_material_ext": T.MATERIAL_EXT_DTYPE}
    cstructs = {"pbrt_b200_scene_desc": H.SceneDesc, "pbrt_b200_render_desc": H.RenderDesc, "pbrt_b200_render_stats": H.RenderStats,
                "pbrt_b200_camera": H.CameraDesc, "pbrt_b200_film": H.FilmDesc, "pbrt_b200_sampler": H.SamplerDesc, "pbrt_b200_integrator": H.IntegratorDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for name, dt in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for field in dt.names:
            lines.append(f'printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    for name, cs in cstructs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cs._fields_:
            lines.append(f'printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines.append("return 0; }")
    (tmp_path / "abi.c").write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(tmp_path / "abi"), str(tmp_path / "abi.c")], check=True)
    out = dict(l.split() for l in subprocess.run([str(tmp_path / "abi")], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, dt in structs.items():
        assert int(out[name]) == dt.itemsize, name
        for field in dt.names:
            assert int(out[f"{name}.{field}"]) == dt.fields[field][1], (name, field)
    for name, cs in cstructs.items():
        assert int(out[name]) == C.sizeof(cs), name
        for field, _ in cs._fields_:
            assert int(out[f"{name}.{field}"]) == getattr(cs, field).offset, (name, field)


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


def test_bvh_structure_check_fast_path_and_general_pass_agree(pkg):
    """scene_create checks the LinearBVHNode array with a segmented pre-order scan (four host threads, stitched like matched
    parentheses) and falls back to the general pass for arrays that are trees but not in pre-order.  Valid arrays of every size -- below
    and above the multi-segment threshold -- must pass (rc = NO_DEVICE here, never INVALID), and every mutation must be caught."""
    lib = pkg.load_library()
    H = pkg.host
    rng = np.random.default_rng(11)

    def create(nodes, n_prims):
        fs = H.FlatScene()
        fs.nodes = np.ascontiguousarray(nodes)
        fs.prims = np.zeros(n_prims, H.PRIM_DTYPE)
        fs.prims["shape_kind"] = 1  # spheres: no index buffers needed
        fs.prims["material"] = -1
        fs.prims["area_light"] = -1
        fs.spheres = np.zeros(1, H.SPHERE_DTYPE)
        d = fs.desc()
        out = C.c_void_p()
        rc = lib.pbrt_b200_scene_create(C.byref(d), 0, C.byref(out))
        if rc == 0:
            lib.pbrt_b200_scene_destroy(out)
        return rc, lib.pbrt_b200_last_error().decode()

    ok = (0, 2)  # created (GPU box) or "no CUDA device" (here): the tables passed
    for n in (3, 1000, 70_000):
        c = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
        nodes, order = H.bvh_build(np.concatenate([c - 0.01, c + 0.01], axis=1), 4, "sah")
        rc, msg = create(nodes, n)
        assert rc in ok, (n, rc, msg)
        if n < 1000:
            continue
        interior = np.flatnonzero(nodes["n_prims"] == 0)
        leaves = np.flatnonzero(nodes["n_prims"] != 0)
        for trial in range(6):
            bad = nodes.copy()
            if trial == 0:
                bad[interior[len(interior) // 2]]["offset"] += 1            # second child off by one: two parents / unreachable node
            elif trial == 1:
                bad[interior[-1]]["offset"] = len(nodes) + 5                 # child out of range
            elif trial == 2:
                bad[leaves[len(leaves) // 3]]["offset"] = n                  # leaf refers past the primitive table
            elif trial == 3:
                bad[interior[len(interior) // 3]]["axis"] = 3
            elif trial == 4:
                bad[leaves[len(leaves) // 2]]["n_prims"] = 0                 # a leaf turned into an interior node with a stale offset
                bad[leaves[len(leaves) // 2]]["offset"] = 1
            else:
                bad = bad[:-1]                                               # the array ends inside a sub-tree
            rc, msg = create(bad, n)
            assert rc == 1, (n, trial, rc, msg)
    # a valid tree that is NOT in pre-order (sub-tree of node 1 = {1, 2, 4}): accepted through the general pass
    t = np.zeros(5, H.NODE_DTYPE)
    t[0]["offset"], t[1]["offset"] = 3, 4
    for leaf, first in ((2, 0), (3, 1), (4, 2)):
        t[leaf]["n_prims"], t[leaf]["offset"] = 1, first
    rc, msg = create(t, 3)
    assert rc in ok, (rc, msg)

    # depth: the traversal stacks one far child per level, 64 entries (bvh.rs:722).  A right-leaning chain keeps the number of PENDING second
    # children at one, so a check that looked at the stack height instead of the depth would pass it
    def chain(depth, tail_nodes=None):
        rows = []
        for k in range(depth):  # interior k: first child = leaf 2k+1, second child = interior 2k+2
            rows.append((0, 2 * k + 2))
            rows.append((1, 0))
        rows.append((1, 0))
        a = np.zeros(len(rows), H.NODE_DTYPE)
        for i, (np_, off) in enumerate(rows):
            a[i]["n_prims"], a[i]["offset"] = np_, off
        return a

    assert create(chain(63), 1)[0] in ok
    rc, msg = create(chain(64), 1)
    assert rc == 1 and "deeper" in msg, (rc, msg)
    # ... and the same two chains hanging off the far end of a big balanced tree (several scan segments)
    c = rng.uniform(-10, 10, (70_000, 3)).astype(np.float32)
    big, _ = H.bvh_build(np.concatenate([c - 0.01, c + 0.01], axis=1), 4, "sah")
    last_leaf = len(big) - 1
    assert big[last_leaf]["n_prims"] != 0
    d0 = 0  # depth of the last leaf = number of interior nodes whose sub-tree contains it
    i = 0
    while big[i]["n_prims"] == 0:
        d0 += 1
        i = int(big[i]["offset"])  # the last node is always reached through second children
    for extra, want_ok in ((63 - d0, True), (64 - d0, False)):
        tail = chain(extra)
        tail["offset"][tail["n_prims"] == 0] += last_leaf
        both = np.concatenate([big[:last_leaf], tail])
        rc, msg = create(both, 70_000)
        assert (rc in ok) == want_ok, (extra, rc, msg)


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
