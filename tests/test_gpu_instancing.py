"""Object instancing (TransformedPrimitive, src/core/primitive.rs:41-103; ObjectBegin/End/Instance, src/core/api.rs:1593-1713).

CPU: the oracle's two-level traversal agrees with the same geometry baked into world space (an independent path through
the oracle).  GPU: hits through instances are bit-identical to the oracle's, and the image passes the relMSE gate."""
import numpy as np
import pytest

REL_MSE_TOL = 1e-3


def test_oracle_instances_agree_with_baked_geometry(pkg, oracle):
    S = pkg.scenes
    inst, baked = S.instanced_scene(), S.instanced_scene(baked=True)
    assert len(inst.flat.instances) == 36 and len(inst.flat.objects) == 3
    assert inst.flat.objects["n_nodes"][1] == 0 and inst.flat.objects["n_prims"][1] == 1  # one-primitive object: no accelerator
    rays = np.concatenate([S.rays_camera(inst.make_integrator(res=(96, 72))), S.rays_diffuse(inst.flat, 20000, seed=3)])
    ha, _ = oracle.intersect(inst.flat, rays)
    hb, _ = oracle.intersect(baked.flat, rays)
    hit_a, hit_b = ha["prim"] != 0xFFFFFFFF, hb["prim"] != 0xFFFFFFFF
    assert np.mean(hit_a == hit_b) > 0.9995  # rays grazing an edge may differ: the transformed ray is not the same float ray
    both = hit_a & hit_b
    # t differs by the transform_ray origin nudge dt (transform.rs:552-556: t_max -= dt, and r.t_max = ray.t_max on a hit)
    assert np.max(np.abs(ha["t"][both] - hb["t"][both]) / hb["t"][both]) < 5e-4
    occ_a, _ = oracle.intersect_p(inst.flat, rays)
    occ_b, _ = oracle.intersect_p(baked.flat, rays)
    assert np.mean(occ_a == occ_b) > 0.9995
    ia, _ = oracle.render_image(inst.flat, inst.make_integrator(spp_=8, res=(64, 48)))
    ib, _ = oracle.render_image(baked.flat, baked.make_integrator(spp_=8, res=(64, 48)))
    assert oracle.rel_mse(ia, ib) < 1e-4


def test_builder_rejects_area_lights_and_nesting_in_objects(pkg):
    H = pkg.host
    b = H.SceneBuilder()
    b.object_begin("a")
    with pytest.raises(H.B200Error):
        b.object_begin("b")
    b.area_light_source("diffuse", L=(1, 1, 1))
    with pytest.raises(H.B200Error):
        b.shape("trianglemesh", P=np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0)], np.float32), indices=np.array([(0, 1, 2)], np.uint32))
    with pytest.raises(H.B200Error):
        b.object_instance("a")
    b.object_end()
    with pytest.raises(H.B200Error):
        b.object_instance("missing")


@pytest.mark.gpu
def test_instanced_hits_bit_exact(pkg, oracle, gpu_lib):
    S = pkg.scenes
    setup = S.instanced_scene()
    rays = np.concatenate([S.rays_camera(setup.make_integrator(res=(320, 240))), S.rays_diffuse(setup.flat, 200000, seed=5)])
    want, _ = oracle.intersect(setup.flat, rays)
    sc = pkg.Scene(setup.flat)
    got = sc.intersect(rays)
    shadow = S.rays_diffuse(setup.flat, 100000, seed=9)
    shadow["t_max"] = 6.0
    occ = sc.intersect_p(shadow)
    sc.close()
    assert (want["prim"] != 0xFFFFFFFF).mean() > 0.3
    assert got.tobytes() == want.tobytes()
    occ_want, _ = oracle.intersect_p(setup.flat, shadow)
    assert np.array_equal(occ, occ_want)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(spp_=16, res=(128, 96)), dict(spp_=8, res=(96, 72), strategy="spatial", sampler_="halton")])
def test_instanced_image_matches_oracle(pkg, oracle, gpu_lib, kw):
    setup = pkg.scenes.instanced_scene()
    integ = setup.make_integrator(**kw)
    sc = pkg.Scene(setup.flat)
    img, stats = integ.render(sc)
    sc.close()
    ref, ostats = oracle.render_image(setup.flat, integ)
    err = oracle.rel_mse(img, ref)
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"
    assert stats.camera_rays == ostats["camera_rays"]
    assert abs(int(stats.intersection_tests) - ostats["intersection_tests"]) <= 0.002 * ostats["intersection_tests"] + 8
