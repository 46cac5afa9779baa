"""Textures, ray differentials, bump mapping, uber / substrate on the GPU against the CPU oracle.

Gate: image relMSE <= 1e-3 (north_star), equal camera-ray counts, ray counts within 0.3 %.  The scene (scenes.textured_scene) holds every
texture kind, every mapping, float and spectrum textures on all seven materials, bump maps on a sphere and on a mesh with per-vertex
normals, and a textured object instance; the second fixture is the reference's own src/scenes/spheres-differentials-texfilt.pbrt
(tests/golden/reference_spheres_scene.pbrt) with a stand-in for the image it names but does not ship
(tests/golden/textures/lines.png, written by tests/golden/make_lines_texture.py)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL_MSE_TOL = 1e-3
GOLDEN = Path(__file__).resolve().parent / "golden"


def _compare(pkg, oracle, flat, integ, tol=REL_MSE_TOL, rays=True):
    sc = pkg.Scene(flat)
    got, st = sc.render(integ)
    sc.close()
    want, ost = oracle.render(flat, integ)
    assert np.isfinite(got).all()
    assert np.array_equal(got[:, 3] > 0, want[:, 3] > 0)
    a, b = oracle.film_resolve(got, integ.film.scale), oracle.film_resolve(want, integ.film.scale)
    err = oracle.rel_mse(a, b)
    assert err <= tol, f"relMSE {err:.3e}"
    assert st.camera_rays == ost["camera_rays"]
    if rays:
        assert abs(int(st.intersection_tests) - ost["intersection_tests"]) <= 0.003 * ost["intersection_tests"] + 8
        assert abs(int(st.shadow_tests) - ost["shadow_tests"]) <= 0.003 * ost["shadow_tests"] + 8
    return err, a, b


@pytest.mark.parametrize("integrator", ["path", "volpath", "whitted", "directlighting:all", "directlighting:one"])
@pytest.mark.parametrize("sampler", ["sobol", "halton"])
def test_textured_scene_matches_the_oracle(pkg, oracle, gpu_lib, integrator, sampler):
    setup = pkg.scenes.textured_scene(xres=128, yres=96, spp=4, sampler=sampler)
    assert len(setup.flat.textures) >= 30 and len(setup.flat.mipmaps) == 3 and setup.flat.materials["textured"].sum() == 9
    _, a, _ = _compare(pkg, oracle, setup.flat, setup.make_integrator(integrator=integrator))
    assert a.mean() > 0.05


@pytest.mark.parametrize("integrator", ["whitted", "directlighting:all"])
def test_textured_scene_under_the_tile_serial_sampler(pkg, oracle, gpu_lib, integrator):
    # the recursive integrators draw a fixed number of samples per surface: the (0,2)-sequence streams stay in step
    setup = pkg.scenes.textured_scene(xres=64, yres=48, spp=2, sampler="02sequence")
    _compare(pkg, oracle, setup.flat, setup.make_integrator(integrator=integrator))


def test_textured_path_under_the_tile_serial_sampler_is_the_same_estimate(pkg, oracle, gpu_lib):
    """PathIntegrator + 02sequence: a tile's PCG32 is threaded through every draw of every path (integrator.rs:302-303), and Russian
    roulette draws a number only when max(beta * etascale) < 1 -- exactly 1 up to rounding along a glass path.  One ulp of difference in a
    texture value (CUDA's atan2f / acosf / expf are not glibc's) flips that test somewhere in a tile and shifts the stream of every later
    sample of the tile: the two images are then different draws of the same estimator.  Checked as such: most tiles agree to rounding,
    the rest agree in the mean."""
    setup = pkg.scenes.textured_scene(xres=128, yres=96, spp=4, sampler="02sequence")
    integ = setup.make_integrator(integrator="path")
    sc = pkg.Scene(setup.flat)
    got, st = sc.render(integ)
    sc.close()
    want, ost = oracle.render(setup.flat, integ)
    a = oracle.film_resolve(got, 1.0).reshape(96, 128, 3)
    b = oracle.film_resolve(want, 1.0).reshape(96, 128, 3)
    assert st.camera_rays == ost["camera_rays"] and np.isfinite(a).all()
    rel = np.abs(a - b).max(axis=2) / (np.abs(b).max(axis=2) + 1e-2)
    tiles = rel.reshape(6, 16, 8, 16).max(axis=(1, 3))
    assert (tiles < 1e-2).mean() >= 0.6, tiles  # tiles whose streams never diverged
    assert np.allclose(a.mean(axis=(0, 1)), b.mean(axis=(0, 1)), rtol=0.03)
    # and with one bounce nothing can diverge (no roulette, no specular chains): the image is the oracle's
    integ1 = setup.make_integrator(integrator="path", maxdepth_=1)
    _compare(pkg, oracle, setup.flat, integ1)


def test_depth_of_field_filters_and_a_baked_instance(pkg, oracle, gpu_lib):
    setup = pkg.scenes.textured_scene(xres=96, yres=72, spp=4, instanced=False)
    _compare(pkg, oracle, setup.flat, setup.make_integrator(lensradius=0.05))  # ray differentials through the lens (perspective.rs:148-165)
    _compare(pkg, oracle, setup.flat, setup.make_integrator(integrator="whitted", lensradius=0.05, filt="gaussian"))
    inst = pkg.scenes.textured_scene(xres=96, yres=72, spp=4, instanced=True)
    sc = pkg.Scene(inst.flat)
    a, _ = sc.render(inst.make_integrator(integrator="whitted"))
    sc.close()
    sc = pkg.Scene(setup.flat)
    b, _ = sc.render(setup.make_integrator(integrator="whitted"))
    sc.close()
    # the first crate is the same geometry instanced / baked: the images differ only where the second (instance-only) crate shows
    d = np.abs(oracle.film_resolve(a, 1.0) - oracle.film_resolve(b, 1.0)).max(axis=1).reshape(72, 96)
    assert (d < 1e-3).mean() > 0.85


def test_textured_scene_at_bench_resolution(pkg, oracle, gpu_lib):
    """T1 at 1920x1080 (what `bench.py --scene t1` times): a strip of 64 tiles through the image centre at 4 spp against the oracle --
    camera-ray footprints at this resolution select other MIP levels than the small renders above."""
    setup = pkg.scenes.textured_scene(xres=1920, yres=1080, spp=4)
    integ = setup.make_integrator()
    nt = integ.n_tiles()
    crop = (nt // 2 - 32, nt // 2 + 32)
    sc = pkg.Scene(setup.flat)
    got, st = sc.render(integ, tile_range=crop)
    sc.close()
    want, ost = oracle.render(setup.flat, integ, tile_range=crop)
    m = want[:, 3] > 0
    assert m.sum() >= 60 * 256 and np.array_equal(got[:, 3] > 0, m) and st.camera_rays == ost["camera_rays"]
    err = oracle.rel_mse(oracle.film_resolve(got[m], 1.0), oracle.film_resolve(want[m], 1.0))
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"


def test_the_reference_texture_scene_renders_like_the_oracle(pkg, oracle, gpu_lib):
    """src/scenes/spheres-differentials-texfilt.pbrt: directlighting, (0,2)-sequence sampler, an EWA-filtered image map on the ground seen
    directly, in a mirror sphere and through a glass sphere (specular_reflect / specular_transmit differentials, integrator.rs:409-520)."""
    api = pkg.pbrt_parse(GOLDEN / "reference_spheres_scene.pbrt", quick_render=True)
    assert not api.errors
    job = api.jobs[0]
    flat = job.flat
    assert len(flat.mipmaps) == 1 and (flat.mipmaps[0]["width"], flat.mipmaps[0]["height"]) == (128, 64)  # 96 x 64 resampled to a power of two
    assert flat.materials["textured"].tolist() == [1, 0, 0]
    err, a, b = _compare(pkg, oracle, flat, job.integrator)
    # the reflection in the mirror sphere is FILTERED: without the specular differentials the lines stay sharp and the images differ by 5e-2
    assert err < 1e-5


def test_textures_on_every_plain_material_slot(pkg, oracle, gpu_lib):
    """Each of the five hot materials with ONE textured parameter at a time (the others constant) against the oracle, plus the
    constant-folded forms: a scale / mix of constants is a constant and the material stays on its fast kernel."""
    from importlib import import_module

    T = import_module("pbrt-rust_b200.textures")
    H = pkg.host
    chk = T.Tex.checkerboard(T.Mapping2D.uv(6.0, 6.0), np.array([0.8, 0.3, 0.2], f32), np.array([0.2, 0.4, 0.8], f32))
    fchk = T.Tex.checkerboard(T.Mapping2D.uv(5.0, 3.0), 0.05, 0.4)
    cases = [("matte", dict(Kd=chk)), ("matte", dict(Kd=0.5, sigma=T.Tex.scale(fchk, 100.0))), ("plastic", dict(Kd=0.3, Ks=chk, roughness=0.2)),
             ("plastic", dict(Kd=chk, roughness=fchk)), ("mirror", dict(Kr=chk)), ("glass", dict(Kt=chk, index=1.4)),
             ("glass", dict(uroughness=fchk, vroughness=0.1)), ("glass", dict(eta=T.Tex.scale(fchk, 4.0))), ("metal", dict(k=chk, roughness=0.1)),
             ("metal", dict(uroughness=fchk, vroughness=0.05)), ("uber", dict(Kd=chk, opacity=T.Tex.scale(chk, 1.2), Kt=0.3, Kr=0.2)),
             ("substrate", dict(Kd=chk, Ks=0.1, uroughness=fchk, vroughness=fchk))]
    for name, kw in cases:
        b = H.SceneBuilder()
        b.light_source("point", **{"from": (2.0, 4.0, 3.0), "I": (25.0, 25.0, 25.0)})
        b.light_source("distant", **{"from": (-1, 3, 2), "to": (0, 0, 0), "L": (1.0, 1.0, 1.0)})
        b.material("matte", Kd=0.6)
        b.shape("trianglemesh", P=[[-6, -1, -6], [6, -1, -6], [6, -1, 6], [-6, -1, 6]], indices=[0, 2, 1, 0, 3, 2])
        b.material(name, **kw)
        b.shape("sphere", radius=1.0)
        flat = b.world_end()
        film = H.Film(64, 64, "box")
        cam = H.PerspectiveCamera(film, H.Transform.look_at((0.5, 1.5, 4.0), (0, 0, 0), (0, 1, 0)).inverse(), fov=40.0)
        for integ in (H.PathIntegrator(cam, film, H.Sampler("sobol", 4), maxdepth=4, lightsamplestrategy="uniform"),
                      H.WhittedIntegrator(cam, film, H.Sampler("halton", 2), maxdepth=4)):
            _compare(pkg, oracle, flat, integ)
    b = H.SceneBuilder()
    b.material("plastic", Kd=np.array([0.2, 0.3, 0.4], f32), roughness=0.3)
    assert not isinstance(b._material, H.TexturedMaterial)


f32 = np.float32
