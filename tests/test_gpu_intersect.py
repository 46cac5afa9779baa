"""Gate 1 (BASELINE.json): for identical ray batches the CUDA closest-hit primitive IDs are
bit-exact vs the oracle and t / barycentrics agree within 1e-5 relative (in practice: bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star tolerance for t and barycentrics


def _compare(pkg, oracle, flat, rays, shadow_rays):
    sc = pkg.Scene(flat)
    got = sc.intersect(rays)
    want, cnt = oracle.intersect(flat, rays)
    assert np.array_equal(got["prim"], want["prim"]), f"{(got['prim'] != want['prim']).sum()} primitive IDs differ"
    for k in ("t", "b0", "b1"):
        a, b = got[k], want[k]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        assert np.all(np.abs(a[fin] - b[fin]) <= REL_TOL * np.abs(b[fin])), k
        # stronger than the gate: the arithmetic is the same op-for-op
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{k} not bit-identical"
    occ = sc.intersect_p(shadow_rays)
    occ_want, _ = oracle.intersect_p(flat, shadow_rays)
    assert np.array_equal(occ, occ_want)
    sc.close()
    return cnt


@pytest.mark.parametrize("which", ["mixed", "cornell", "spheres", "sphere65k"])
def test_hits_bit_exact(pkg, oracle, gpu_lib, which):
    S = pkg.scenes
    setup = {"mixed": S.small_mixed_scene, "cornell": S.cornell_scene, "spheres": S.spheres_scene,
             "sphere65k": lambda: S.displaced_sphere_scene(256, 128)}[which]()
    flat = setup.flat
    n = 300_000
    rays = np.concatenate([S.rays_diffuse(flat, n, seed=7), S.rays_camera(setup.make_integrator(), max_rays=100_000)])
    # axis-parallel and zero-component directions exercise the inf/NaN slab paths
    special = S.rays_diffuse(flat, 30_000, seed=3)
    special["d"][:10_000, 0] = 0.0
    special["d"][10_000:20_000, 1] = 0.0
    special["d"][20_000:, 2] = -0.0
    rays = np.concatenate([rays, special])
    shadow = S.rays_shadow(flat, n, seed=11) if len(flat.tri_indices) else rays[:1000]
    cnt = _compare(pkg, oracle, flat, rays, np.concatenate([shadow, rays[:50_000]]))
    assert cnt[2] == len(rays)


def test_empty_and_ragged(pkg, oracle, gpu_lib):
    S = pkg.scenes
    flat = S.small_mixed_scene().flat
    sc = pkg.Scene(flat)
    assert len(sc.intersect(np.zeros(0, pkg.host.RAY_DTYPE))) == 0
    for n in (1, 31, 33, 127, 129, 1000):
        rays = S.rays_diffuse(flat, n, seed=n)
        want, _ = oracle.intersect(flat, rays)
        got = sc.intersect(rays)
        assert got.tobytes() == want.tobytes()
    # t_max clipping: a ray that stops short of the first surface reports a miss
    rays = S.rays_diffuse(flat, 5000, seed=99)
    full, _ = oracle.intersect(flat, rays)
    rays["t_max"] = np.where(full["prim"] != pkg.host.NO_HIT, full["t"] * 0.5, 1.0).astype(np.float32)
    want, _ = oracle.intersect(flat, rays)
    got = sc.intersect(rays)
    assert got.tobytes() == want.tobytes()
    sc.close()
    # empty scene: every ray misses
    empty = pkg.SceneBuilder().world_end()
    sc = pkg.Scene(empty)
    got = sc.intersect(S.rays_diffuse(flat, 100))
    assert np.all(got["prim"] == pkg.host.NO_HIT)
    sc.close()
