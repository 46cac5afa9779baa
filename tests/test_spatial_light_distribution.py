"""SpatialLightDistribution (src/core/lightdistrib.rs:105-340), the reference's DEFAULT `lightsamplestrategy`.

CPU: properties of the oracle's restatement.  GPU: the device tables (eager and lazy builds) match the oracle's
per-voxel `light_contrib` to float tolerance (the values go through sqrt/normalize chains that are IEEE-exact and, for
infinite lights only, sin/cos from different libms), and the rendered image passes the relMSE gate."""
import numpy as np
import pytest

REL_MSE_TOL = 1e-3


def _points(flat, n, seed=3):
    wb = np.asarray(flat.world_bound, np.float32).reshape(-1)
    u = np.random.default_rng(seed).random((n, 3), dtype=np.float32)
    return (wb[:3] + (wb[3:] - wb[:3]) * u).astype(np.float32)


def test_oracle_voxel_grid_and_floor(pkg, oracle):
    setup = pkg.scenes.many_lights_scene()
    flat = setup.flat
    pts = _points(flat, 300)
    voxel, func, nv = oracle.spatial_lookup(flat, pts)
    wb = np.asarray(flat.world_bound, np.float64).reshape(-1)
    diag = wb[3:] - wb[:3]
    # lightdistrib.rs:113-130: 64 voxels along the widest axis, roughly cubic voxels elsewhere
    assert nv.max() == 64 and nv[int(np.argmax(diag))] == 64
    assert np.all(nv == np.maximum(1, np.round(diag / diag.max() * 64)).astype(int))
    assert np.all(voxel >= 0) and np.all(voxel < nv)
    expect = np.clip(np.floor((pts - wb[:3]) / diag * nv), 0, nv - 1).astype(int)
    assert np.mean(np.all(voxel == expect, axis=1)) > 0.99  # f32 vs f64 boundary cases only
    # lightdistrib.rs:205-219: no light below 0.1% of the average contribution
    assert func.shape[1] == len(flat.lights) and np.all(func > 0)
    avg = func.sum(1, keepdims=True) / (128.0 * func.shape[1])  # light_contrib holds sums over the 128 points
    assert np.all(func >= 0.001 * avg * 0.9)
    # same voxel => same distribution
    _, f2, _ = oracle.spatial_lookup(flat, pts[:50])
    assert f2.tobytes() == func[:50].tobytes()


def test_oracle_single_light_falls_back_to_uniform(pkg, oracle):
    """create_light_sample_distribution: one light => UniformLightDistribution whatever the name (lightdistrib.rs:21)."""
    setup = pkg.scenes.spheres_scene()
    a, _ = oracle.render(setup.flat, setup.make_integrator(spp_=2, res=(32, 32), strategy="spatial"))
    b, _ = oracle.render(setup.flat, setup.make_integrator(spp_=2, res=(32, 32), strategy="uniform"))
    assert a.tobytes() == b.tobytes()


def test_oracle_spatial_differs_from_power_but_agrees_in_expectation(pkg, oracle):
    setup = pkg.scenes.many_lights_scene()
    kw = dict(spp_=64, res=(48, 27))
    a, _ = oracle.render_image(setup.flat, setup.make_integrator(strategy="spatial", **kw))
    b, _ = oracle.render_image(setup.flat, setup.make_integrator(strategy="power", **kw))
    assert a.tobytes() != b.tobytes()
    assert abs(a.mean() - b.mean()) < 0.05 * b.mean()  # both are unbiased estimators of the same image


@pytest.mark.gpu
@pytest.mark.parametrize("lazy", [False, True])
@pytest.mark.parametrize("scene", ["many_lights_scene", "small_mixed_scene", "cornell_scene"])
def test_device_tables_match_oracle(pkg, oracle, gpu_lib, scene, lazy):
    setup = getattr(pkg.scenes, scene)()
    flat = setup.flat
    pts = _points(flat, 4000)
    want_v, want_f, nv = oracle.spatial_lookup(flat, pts)
    sc = pkg.Scene(flat)
    got_v, got_f = sc.light_distribution_lookup(pts, "spatial", lazy=lazy)
    if lazy:  # a second lookup finds the voxels already built
        again_v, again_f = sc.light_distribution_lookup(pts, "spatial", lazy=True)
        assert again_f.tobytes() == got_f.tobytes() and again_v.tobytes() == got_v.tobytes()
    sc.close()
    assert np.array_equal(got_v, want_v)
    assert np.all(got_f > 0)
    np.testing.assert_allclose(got_f, want_f, rtol=2e-5)
    has_infinite = bool(np.any(flat.lights["type"] == 4))
    if not has_infinite:  # everything else is +,-,*,/,sqrt in the reference's order => bit-identical
        assert got_f.tobytes() == want_f.tobytes()


@pytest.mark.gpu
def test_uniform_and_power_lookup(pkg, oracle, gpu_lib):
    setup = pkg.scenes.small_mixed_scene()
    sc = pkg.Scene(setup.flat)
    pts = _points(setup.flat, 16)
    v, f = sc.light_distribution_lookup(pts, "uniform")
    assert np.all(v == -1) and np.all(f == 1.0)
    v, f = sc.light_distribution_lookup(pts, "power")
    sc.close()
    assert np.all(v == -1) and np.all(f == f[0]) and len(np.unique(f[0])) > 1


@pytest.mark.gpu
@pytest.mark.parametrize("lazy", [False, True])
@pytest.mark.parametrize("scene,kw", [("many_lights_scene", dict(spp_=16, res=(128, 72))), ("small_mixed_scene", dict(spp_=16, res=(96, 64), strategy="spatial")),
                                      ("cornell_scene", dict(spp_=8, res=(96, 96), strategy="spatial", filt="box"))])
def test_spatial_image_matches_oracle(pkg, oracle, gpu_lib, scene, kw, lazy):
    setup = getattr(pkg.scenes, scene)()
    integ = setup.make_integrator(**kw)
    assert integ.desc().integrator.light_sample_strategy == pkg.host.LIGHTS_SPATIAL
    sc = pkg.Scene(setup.flat)
    img, stats = integ.render(sc, flags=pkg.host.RENDER_LAZY_SPATIAL if lazy else 0)
    sc.close()
    ref, ostats = oracle.render_image(setup.flat, integ)
    err = oracle.rel_mse(img, ref)
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"
    assert stats.camera_rays == ostats["camera_rays"]
    assert abs(int(stats.shadow_tests) - ostats["shadow_tests"]) <= 0.002 * ostats["shadow_tests"] + 8
