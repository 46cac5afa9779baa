"""Independent check of the scene FLATTENING (host.py's material / light / primitive rows).

Every GPU-vs-oracle parity test feeds both sides the same flattened tables, so a wrong row (a material default, a light's
area, `two_sided`, the order of `scene.lights`) would be invisible to them.  Here the expected rows are written down by hand
from the reference's own creation code -- the literals below are the defaults and formulas of

  materials/matte.rs:56-57, plastic.rs:73-77, mirror.rs:45, glass.rs:96-105, metal.rs:116-122 (copper: spectrum.rs COPPER_N / COPPER_K),
  lights/point.rs:99-106, distant.rs:124-132, spot.rs:119-146, diffuse.rs:178-195 (+ api.rs:1531-1546: one light per triangle,
  created when the shape is, in file order), api.rs `reverse_orientation` / `swaps_handedness` flags

-- and compared with what `SceneBuilder` produced, without going through any helper of host.py.
"""
import math

import numpy as np

F = np.float32


def _tri_area64(p):
    return 0.5 * float(np.linalg.norm(np.cross(p[1].astype(np.float64) - p[0].astype(np.float64), p[2].astype(np.float64) - p[0].astype(np.float64))))


def test_rows_match_hand_derived_expectations(pkg):
    H = pkg.host
    b = H.SceneBuilder()
    # --- file order: point light, then shapes (two of them emissive), then distant + spot lights
    b.light_source("point", **{"from": (1.0, 2.0, 1.0), "I": (2.0, 4.0, 6.0), "scale": 0.5})
    b.material("matte")                                                          # all defaults
    b.shape("trianglemesh", P=[[0, 0, 0], [2, 0, 0], [0, 3, 0]], indices=[0, 1, 2])           # prim 0
    b.material("plastic")
    b.attribute_begin()
    b.area_light_source("diffuse", L=(3.0, 2.0, 1.0), scale=2.0, twosided=True)
    quad = np.array([[0, 0, 5], [1, 0, 5], [1, 2, 5], [0, 2, 5]], np.float32)
    b.shape("trianglemesh", P=quad, indices=[0, 1, 2, 0, 2, 3])                  # prims 1, 2 -> two area lights
    b.attribute_end()
    b.material("mirror")
    b.attribute_begin()
    b.scale(1.0, 1.0, -1.0)                                                      # a reflection: swaps handedness
    b.reverse_orientation = True
    b.shape("trianglemesh", P=[[0, 0, 1], [1, 0, 1], [0, 1, 1]], indices=[0, 1, 2], N=[[0, 0, 1]] * 3, uv=[[0, 0], [1, 0], [0, 1]])  # prim 3
    b.attribute_end()
    b.material("glass")
    b.shape("trianglemesh", P=[[5, 0, 0], [6, 0, 0], [5, 1, 0]], indices=[0, 1, 2])           # prim 4
    b.material("metal")
    b.shape("trianglemesh", P=[[7, 0, 0], [8, 0, 0], [7, 1, 0]], indices=[0, 1, 2])           # prim 5
    b.material("glass", index=1.33, uroughness=0.2, vroughness=0.3, remaproughness=False, Kr=(0.9, 0.8, 0.7))
    b.shape("trianglemesh", P=[[9, 0, 0], [10, 0, 0], [9, 1, 0]], indices=[0, 1, 2])          # prim 6
    b.light_source("distant", **{"from": (0.0, 0.0, 10.0), "to": (0.0, 0.0, 0.0), "L": (1.0, 1.0, 2.0)})
    b.light_source("spot", **{"from": (0.0, 5.0, 0.0), "to": (0.0, 0.0, 0.0), "I": 10.0, "coneangle": 40.0, "conedeltaangle": 10.0})
    flat = b.world_end()

    # ---- materials, in order of first use
    m = flat.materials
    assert len(m) == 6
    copper_n = np.array([0.19999069, 0.92208463, 1.09987593], np.float32)  # Spectrum::from_sampled(COPPER_WAVELENGTHS, COPPER_N): checked in
    copper_k = np.array([3.90463543, 2.44763327, 2.13765264], np.float32)  # tests/test_scene_file_frontend.py against the CIE tables
    want = [
        dict(type=0, a=(0.5, 0.5, 0.5), b=(0, 0, 0), f0=0.0, f1=0.0, f2=0.0, remap=1),                 # matte: Kd 0.5, sigma 0
        dict(type=1, a=(0.25,) * 3, b=(0.25,) * 3, f0=0.1, f1=0.0, f2=0.0, remap=1),                     # plastic: Kd .25 Ks .25 roughness .1 remap
        dict(type=2, a=(0.9,) * 3, b=(0, 0, 0), f0=0.0, f1=0.0, f2=0.0, remap=1),                        # mirror: Kr 0.9
        dict(type=3, a=(1.0,) * 3, b=(1.0,) * 3, f0=0.0, f1=0.0, f2=1.5, remap=1),                       # glass: Kr 1 Kt 1 index 1.5 rough 0
        dict(type=4, a=tuple(copper_n), b=tuple(copper_k), f0=0.01, f1=0.01, f2=0.0, remap=1),          # metal: copper, roughness .01 both axes
        dict(type=3, a=(0.9, 0.8, 0.7), b=(1.0,) * 3, f0=0.2, f1=0.3, f2=1.33, remap=0),
    ]
    for row, w in zip(m, want):
        assert int(row["type"]) == w["type"] and int(row["remap_roughness"]) == w["remap"], (row, w)
        assert np.allclose(row["a"], w["a"], rtol=2e-6) and np.allclose(row["b"], w["b"], rtol=2e-6), (row, w)
        assert np.allclose([row["f0"], row["f1"], row["f2"]], [w["f0"], w["f1"], w["f2"]], rtol=1e-6), (row, w)

    # ---- lights: scene.lights order = creation order (point, the two triangles of the emissive quad, distant, spot)
    L = flat.lights
    assert [int(t) for t in L["type"]] == [H.LIGHT_POINT, H.LIGHT_DIFFUSE, H.LIGHT_DIFFUSE, H.LIGHT_DISTANT, H.LIGHT_SPOT]
    # point: I * scale at translate(P.x, P.y, P.x) -- the reference's typo (point.rs:103), kept
    assert np.allclose(L[0]["L"], (1.0, 2.0, 3.0)) and np.allclose(L[0]["pos"], (1.0, 2.0, 1.0))
    tris = quad[[0, 1, 2]], quad[[0, 2, 3]]
    for k in (1, 2):
        assert np.allclose(L[k]["L"], (6.0, 4.0, 2.0)) and int(L[k]["two_sided"]) == 1 and int(L[k]["n_samples"]) == 1
        assert int(L[k]["shape_kind"]) == H.SHAPE_TRIANGLE
        tri = flat.vertex_p[flat.tri_indices[int(L[k]["shape_index"])]]
        assert np.array_equal(tri, tris[k - 1])
        assert abs(float(L[k]["area"]) - _tri_area64(tris[k - 1])) <= 1e-6 * _tri_area64(tris[k - 1])  # 1.0 each
    d = np.array([0.0, 0.0, 10.0]) / 10.0  # normalize(from - to)
    assert np.allclose(L[3]["dir"], d) and np.allclose(L[3]["L"], (1.0, 1.0, 2.0))
    assert np.allclose(L[4]["L"], (10.0,) * 3) and np.allclose(L[4]["pos"], (0.0, 5.0, 0.0), atol=1e-6)
    assert abs(float(L[4]["cos_total_width"]) - math.cos(math.radians(40.0))) < 1e-6
    assert abs(float(L[4]["cos_falloff_start"]) - math.cos(math.radians(30.0))) < 1e-6
    # world_to_light takes the spot's axis (0, -1, 0) to +z
    w2l = L[4]["world_to_light"].reshape(4, 4).astype(np.float64)
    axis = w2l[:3, :3] @ np.array([0.0, -1.0, 0.0])
    assert np.allclose(axis, (0.0, 0.0, 1.0), atol=1e-6)
    assert np.allclose(w2l @ np.array([0.0, 5.0, 0.0, 1.0]), (0.0, 0.0, 0.0, 1.0), atol=1e-5)

    # ---- primitive rows (BVH slot order -> look them up by creation index)
    pr = flat.prims[np.argsort(flat.prims["creation_index"])]
    assert len(pr) == 7
    assert [int(x) for x in pr["material"]] == [0, 1, 1, 2, 3, 4, 5]
    assert [int(x) for x in pr["area_light"]] == [-1, 1, 2, -1, -1, -1, -1]
    assert int(L[1]["shape_index"]) == int(pr[1]["shape_index"]) and int(L[2]["shape_index"]) == int(pr[2]["shape_index"])
    fl = [int(x) for x in pr["flags"]]
    assert fl[0] == 0 and fl[1] == 0 and fl[4] == 0
    assert fl[3] == (H.PRIM_REVERSE_ORIENTATION | H.PRIM_SWAPS_HANDEDNESS | H.PRIM_HAS_N | H.PRIM_HAS_UV)
    # the mirrored triangle's vertices and normals went through the CTM (z -> -z; normals by the inverse transpose)
    t3 = flat.vertex_p[flat.tri_indices[int(pr[3]["shape_index"])]]
    assert np.allclose(t3[:, 2], -1.0)
    n3 = flat.vertex_n[flat.tri_indices[int(pr[3]["shape_index"])]]
    assert np.allclose(n3, (0.0, 0.0, -1.0))
    # world bound = union of everything (Scene.wb)
    wb = flat.nodes[0]["bounds"]
    assert np.allclose(wb, (0, 0, -1, 10, 3, 5))
