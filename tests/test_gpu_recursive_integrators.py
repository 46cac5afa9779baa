"""DirectLightingIntegrator ("one" / "all") and WhittedIntegrator on the device vs the CPU oracle's
restatement of directlighting.rs / whitted.rs / integrator.rs:40-79,409-520.  Same north-star image gate as the path
integrator: relMSE <= 1e-3 against the oracle's render with the same sampler, same camera-ray count."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL_MSE_TOL = 1e-3


def _make(pkg, setup, kind, spp, res, sampler="sobol", maxdepth=5, **kw):
    base = setup.make_integrator(spp_=spp, res=res, sampler_=sampler)
    H = pkg.host
    if kind == "whitted":
        return H.WhittedIntegrator(base.camera, base.film, base.sampler, maxdepth=maxdepth)
    return H.DirectLightingIntegrator(base.camera, base.film, base.sampler, maxdepth=maxdepth, strategy=kind)


CASES = {
    # mirror + glass spheres over a matte quad, one distant light: the reference's own spheres scene; specular recursion both ways
    "spheres-whitted": ("spheres_scene", {}, "whitted", dict(spp=8, res=(120, 120))),
    "spheres-direct-all": ("spheres_scene", {}, "all", dict(spp=8, res=(120, 120))),
    "spheres-direct-one-depth8": ("spheres_scene", {}, "one", dict(spp=4, res=(96, 96), maxdepth=8)),
    # every material, point + distant + area + infinite lights: shadow entries for 5 lights per hit, MIS rays to area / infinite lights
    "mixed-whitted": ("small_mixed_scene", {}, "whitted", dict(spp=8, res=(96, 64))),
    "mixed-direct-all": ("small_mixed_scene", {}, "all", dict(spp=8, res=(96, 64))),
    "mixed-direct-one-halton": ("small_mixed_scene", {}, "one", dict(spp=8, res=(96, 64), sampler="halton")),
    "mixed-direct-all-halton-depth1": ("small_mixed_scene", {}, "all", dict(spp=4, res=(96, 64), sampler="halton", maxdepth=1)),
    # two-triangle area light: estimate_direct's light and BSDF halves with MIS weights
    "cornell-direct-all": ("cornell_scene", {}, "all", dict(spp=8, res=(80, 80))),
    "cornell-direct-one": ("cornell_scene", {}, "one", dict(spp=8, res=(80, 80))),
    # object instances under both integrator families
    "instanced-whitted": ("instanced_scene", {}, "whitted", dict(spp=4, res=(96, 72))),
    "instanced-direct-all": ("instanced_scene", {}, "all", dict(spp=4, res=(96, 72))),
}


@pytest.mark.parametrize("name", list(CASES))
def test_recursive_integrator_matches_oracle(pkg, oracle, gpu_lib, name):
    gen, gkw, kind, kw = CASES[name]
    setup = getattr(pkg.scenes, gen)(**gkw)
    integ = _make(pkg, setup, kind, **kw)
    sc = pkg.Scene(setup.flat)
    img, stats = integ.render(sc)
    sc.close()
    ref, ostats = oracle.render_image(setup.flat, integ)
    assert np.isfinite(img).all()
    err = oracle.rel_mse(img, ref)
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"
    assert stats.camera_rays == ostats["camera_rays"]
    assert abs(int(stats.shadow_tests) - ostats["shadow_tests"]) <= 0.002 * ostats["shadow_tests"] + 8


def test_small_queue_and_tile_windows(pkg, oracle, gpu_lib):
    """Few path slots (many regenerations, stack reuse) and split tile / sample windows give the one-shot image."""
    setup = pkg.scenes.spheres_scene()
    integ = _make(pkg, setup, "whitted", spp=4, res=(64, 64))
    sc = pkg.Scene(setup.flat)
    full, _ = sc.render(integ)
    small, _ = sc.render(integ, paths_in_flight=512)
    nt = integ.n_tiles()
    acc = np.zeros_like(full)
    for tr in ((0, nt // 2), (nt // 2, nt)):
        for sr in ((0, 1), (1, 4)):
            sc.render(integ, rgbw=acc, tile_range=tr, sample_range=sr)
    sc.close()
    assert np.allclose(small, full, rtol=2e-5, atol=2e-5)
    assert np.allclose(acc, full, rtol=2e-5, atol=2e-5)


def test_unsupported_combinations_fail_loudly(pkg, gpu_lib):
    setup = pkg.scenes.many_lights_scene()  # > 50 lights: "all" would need more than the 1024 Sobol' dimensions (the reference panics)
    assert len(setup.flat.lights) > 51
    integ = _make(pkg, setup, "all", spp=4, res=(32, 32))
    sc = pkg.Scene(setup.flat)
    with pytest.raises(pkg.B200Error):
        sc.render(integ)
    sc.close()


@pytest.mark.parametrize("kind,kw", [("whitted", dict(spp=4, res=(70, 50))), ("one", dict(spp=2, res=(64, 48))), ("all", dict(spp=4, res=(64, 48), maxdepth=3))])
def test_02sequence_sampler_under_the_recursive_integrators(pkg, oracle, gpu_lib, kind, kw):
    """Tile-serial (0,2)-sequence stream: get_1d / get_2d beyond the precomputed dimensions draw from the tile's PCG32 in the
    recursion's depth-first order, and the 2D sample arrays of "all" are extra sobol_2d tables filled by start_pixel."""
    setup = pkg.scenes.small_mixed_scene()
    integ = _make(pkg, setup, kind, sampler="02sequence", **kw)
    sc = pkg.Scene(setup.flat)
    img, stats = integ.render(sc)
    sc.close()
    ref, ostats = oracle.render_image(setup.flat, integ)
    err = oracle.rel_mse(img, ref)
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"
    assert stats.camera_rays == ostats["camera_rays"]


def test_reference_spheres_scene_file(pkg, oracle, gpu_lib, tmp_path):
    """tests/golden/reference_spheres_scene.pbrt holds the directives and parameters of the reference's
    src/scenes/spheres-differentials-texfilt.pbrt (re-emitted from the parsed command list; tests/test_scene_file_frontend.py
    checks that both files parse to the same commands whenever the reference tree is mounted):
    Integrator "directlighting" maxdepth 10 (strategy "all"), Sampler "lowdiscrepancy" 1 spp, 1000x500, a checkerboard texture
    nobody uses and an image map whose file does not exist in the reference tree either (-> the constant grey texture of
    imagemap.rs:136-142).  Parsed, flattened and rendered on the device; compared with the oracle's render of the same job."""
    import shutil
    from pathlib import Path
    import warnings
    shutil.copy(Path(__file__).parent / "golden" / "reference_spheres_scene.pbrt", tmp_path / "spheres.pbrt")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        job = pkg.pbrt_parse(tmp_path / "spheres.pbrt").jobs[0]
    assert (job.integrator.name, job.integrator.strategy, job.integrator.max_depth) == ("directlighting", "all", 10)
    assert job.sampler.kind == pkg.host.SAMPLER_ZEROTWO and job.sampler.spp == 1 and job.film.full_resolution == (1000, 500)
    assert abs(float(job.flat.materials[0]["a"][0]) - 0.21404114) < 1e-7  # inverse_gamma_correct(0.5): lines.png is missing
    img, stats = job.render(device=0)
    ref, ostats = oracle.render_image(job.flat, job.integrator)
    err = oracle.rel_mse(img, ref)
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"
    assert stats.camera_rays == ostats["camera_rays"] == 500000


def test_scene_file_with_directlighting(pkg, oracle, gpu_lib, tmp_path):
    """The reference's src/scenes/spheres-differentials-texfilt.pbrt, verbatim except for the image-map texture (constant Kd)
    and the sampler (sobol instead of lowdiscrepancy): Integrator "directlighting" "integer maxdepth" [10]."""
    text = '''LookAt 2 2 5   0 -.4 0 0 1 0
Camera "perspective" "float fov" [30 ]
Film "image" "integer xresolution" [200 ] "integer yresolution" [100 ]
Integrator "directlighting" "integer maxdepth" [10]
Sampler "sobol" "integer pixelsamples" [4]
PixelFilter "box"
WorldBegin
LightSource "distant" "point from" [0 10 0 ] "point to" [0 0 0 ] "color L" [3.141593 3.141593 3.141593 ]
AttributeBegin
  Translate .25 0 0
  Material "matte" "color Kd" [.5 .5 .5]
  Shape "trianglemesh"  "integer indices" [0 2 1 0 3 2 ] "point P" [-100 -1 -100 400 -1 -100 400 -1 400 -100 -1 400 ] "float st" [ 0 0 1 0 0 1 1 1]
AttributeEnd
Translate -1.3 0 0
Material "mirror"
Shape "sphere"
Translate 2.6 0 0
Material "glass"
Shape "sphere"
WorldEnd
'''
    job = pkg.pbrt_parse_string(text).jobs[0]
    assert job.integrator.name == "directlighting" and job.integrator.strategy == "all" and job.integrator.max_depth == 10
    img, stats = job.render(device=0)
    ref, ostats = oracle.render_image(job.flat, job.integrator)
    assert oracle.rel_mse(img, ref) <= REL_MSE_TOL and stats.camera_rays == ostats["camera_rays"]


@pytest.mark.parametrize("sampler", ["sobol", "halton"])
def test_directlighting_all_with_multi_sample_lights(pkg, oracle, gpu_lib, sampler):
    """Lights asking for several samples ("integer samples" n): uniform_sample_all_lights averages n estimate_direct calls fed
    from n-element sample arrays (integrator.rs:63-74); element j of an array is evaluated at get_index_for_sample(j)."""
    H = pkg.host
    b = H.SceneBuilder()
    b.material("matte", Kd=(0.5, 0.5, 0.5))
    P, I = pkg.scenes.quad((-6, -6, 0), (6, -6, 0), (6, 6, 0), (-6, 6, 0))
    b.shape("trianglemesh", P=P, indices=I)
    b.attribute_begin()
    b.translate(-1.0, 0.3, 0.8)
    b.material("plastic", Kd=(0.3, 0.5, 0.2), Ks=0.4, roughness=0.05)
    b.shape("sphere", radius=0.8)
    b.attribute_end()
    b.attribute_begin()
    b.translate(1.2, 0.0, 0.6)
    b.material("glass")
    b.shape("sphere", radius=0.6)
    b.attribute_end()
    b.attribute_begin()
    b.area_light_source("diffuse", L=(12, 11, 9), samples=4)
    Pl, Il = pkg.scenes.quad((-1, -1, 3.5), (-1, 1, 3.5), (1, 1, 3.5), (1, -1, 3.5))
    b.shape("trianglemesh", P=Pl, indices=Il)
    b.attribute_end()
    b.attribute_begin()
    b.translate(2.5, -2.0, 2.0)
    b.area_light_source("diffuse", L=(4, 6, 9), twosided=True, samples=3)
    b.shape("sphere", radius=0.3)
    b.attribute_end()
    b.light_source("infinite", L=(0.1, 0.1, 0.15), samples=2)
    b.light_source("point", I=(3, 3, 3), **{"from": (-3, -3, 3)})
    flat = b.world_end()
    assert sorted(set(flat.lights["n_samples"].tolist())) == [1, 2, 3, 4]
    film = H.Film(80, 56, "box")
    cam = H.PerspectiveCamera(film, H.Transform.look_at((0, -6, 2.5), (0, 0, 0.7), (0, 0, 1)).inverse(), fov=42.0)
    integ = H.DirectLightingIntegrator(cam, film, H.Sampler(sampler, 4), maxdepth=4, strategy="all")
    sc = pkg.Scene(flat)
    img, stats = integ.render(sc)
    sc.close()
    ref, ostats = oracle.render_image(flat, integ)
    err = oracle.rel_mse(img, ref)
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"
    assert stats.camera_rays == ostats["camera_rays"]
    assert abs(int(stats.shadow_tests) - ostats["shadow_tests"]) <= 0.002 * ostats["shadow_tests"] + 8
