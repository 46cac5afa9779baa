"""Sphere area lights in the oracle (Sphere::sample_interaction / pdf_wi, sphere.rs:313-395).  The reference's
own sphere_solid_angle test goes through Sphere::solid_angle's closed form and does not pin the sampling code, so the
restatement is cross-checked through what the estimator must satisfy:

* a two-sided sphere light and a finely tessellated two-sided mesh of the same sphere give the same image in expectation
  (light sampling + BSDF sampling + MIS weights only agree if sample_interaction's density and pdf_wi describe each other);
* the reference's cone branch never sets the sampled point's normal (sphere.rs:371-374), so a ONE-sided sphere light emits
  nothing through light sampling: under the Whitted integrator (light sampling only) it lights nothing but is visible itself.
"""
import importlib

import numpy as np

pkg = importlib.import_module("pbrt-rust_b200")
H = pkg.host


def test_sphere_light_matches_tessellated_light(oracle):
    a = pkg.scenes.sphere_lights_scene(enclosing=False)
    b = pkg.scenes.sphere_lights_scene(enclosing=False, tessellated=True)
    assert len(a.flat.lights) == 2 and a.flat.lights["shape_kind"].tolist() == [H.SHAPE_SPHERE, H.SHAPE_SPHERE]
    assert abs(a.flat.lights["area"][0] - 4 * np.pi * 0.35 ** 2) < 1e-5
    for depth in (1, 3):
        ia, _ = oracle.render_image(a.flat, a.make_integrator(spp_=96, res=(48, 32), maxdepth_=depth, strategy="uniform"))
        ib, _ = oracle.render_image(b.flat, b.make_integrator(spp_=96, res=(48, 32), maxdepth_=depth, strategy="power"))
        assert abs(ia.mean() / ib.mean() - 1.0) < 0.01, depth  # the inscribed 64x32 mesh has 0.2 % less area
        assert oracle.rel_mse(ia, ib) < 2e-3


def test_one_sided_sphere_light_emits_only_when_hit(oracle):
    b = H.SceneBuilder()
    b.material("matte", Kd=0.5)
    P, I = pkg.scenes.quad((-8, -8, 0), (8, -8, 0), (8, 8, 0), (-8, 8, 0))
    b.shape("trianglemesh", P=P, indices=I)
    b.attribute_begin()
    b.translate(0, 0, 2)
    b.area_light_source("diffuse", L=(5, 5, 5))
    b.shape("sphere", radius=0.5)
    b.attribute_end()
    flat = b.world_end()
    film = H.Film(48, 32, "box")
    cam = H.PerspectiveCamera(film, H.Transform.look_at((0, -6, 2.5), (0, 0, 1.0), (0, 0, 1)).inverse(), fov=45.0)
    whitted = H.WhittedIntegrator(cam, film, H.Sampler("sobol", 4), maxdepth=3)
    img, _ = oracle.render_image(flat, whitted)
    lit = img.sum(axis=2) > 0
    # a pixel is (samples that see the light itself) x 5 / (samples in the pixel); everything else -- the whole floor -- stays black
    assert 20 < lit.sum() < 200 and img[lit].max() <= 5.0 + 1e-4 and np.isclose(img.max(), 5.0, atol=1e-4)
    path = H.PathIntegrator(cam, film, H.Sampler("sobol", 16), maxdepth=2, lightsamplestrategy="uniform")
    img2, _ = oracle.render_image(flat, path)
    assert (img2.sum(axis=2) > 0).sum() > 3 * lit.sum()  # BSDF-sampled rays that hit the light do see its emission


def test_sphere_lights_golden(oracle):
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "sphere_lights_golden.npz")
    setup = pkg.scenes.sphere_lights_scene()
    for strategy in ("power", "spatial"):
        img, _ = oracle.render_image(setup.flat, setup.make_integrator(spp_=4, res=(48, 32), strategy=strategy))
        assert np.allclose(img, g[strategy], rtol=1e-5, atol=1e-6), strategy
