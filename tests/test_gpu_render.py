"""Gate 2 (BASELINE.json): the image rendered by the CUDA wavefront path matches the CPU oracle's
render of the same scene with the same sampler within relMSE <= 1e-3."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_MSE_TOL = 1e-3  # north_star: image relMSE <= 1e-3 vs the CPU render


def _render_both(pkg, oracle, setup, **kw):
    integ = setup.make_integrator(**kw)
    sc = pkg.Scene(setup.flat)
    img, stats = integ.render(sc)
    sc.close()
    ref, ostats = oracle.render_image(setup.flat, integ)
    return img, stats, ref, ostats


CASES = {
    "spheres-sobol": (lambda S: S.spheres_scene(), dict(spp_=16, res=(160, 160))),
    "cornell-power-gaussian": (lambda S: S.cornell_scene(), dict(spp_=16, res=(128, 128))),
    "cornell-uniform-box": (lambda S: S.cornell_scene(), dict(spp_=8, res=(96, 96), strategy="uniform", filt="box")),
    "mixed-all-materials": (lambda S: S.small_mixed_scene(), dict(spp_=16, res=(144, 96))),
    "mixed-halton": (lambda S: S.small_mixed_scene(), dict(spp_=8, res=(96, 64), sampler_="halton")),
    "mixed-depth1": (lambda S: S.small_mixed_scene(), dict(spp_=4, res=(96, 64), maxdepth_=1)),
    "mixed-depth0": (lambda S: S.small_mixed_scene(), dict(spp_=4, res=(96, 64), maxdepth_=0)),
    "sphere16k-normals": (lambda S: S.displaced_sphere_scene(128, 64), dict(spp_=8, res=(160, 90))),
    # ZeroTwoSequenceSampler: tile-serial on the device, so the per-tile PCG32 stream is the reference's
    "mixed-02sequence": (lambda S: S.small_mixed_scene(), dict(spp_=8, res=(80, 56), sampler_="02sequence")),
    "cornell-02sequence-spp-not-pow2": (lambda S: S.cornell_scene(), dict(spp_=6, res=(48, 40), sampler_="02sequence", filt="box")),
    "spheres-02sequence-partial-tiles": (lambda S: S.spheres_scene(), dict(spp_=4, res=(70, 50), sampler_="02sequence")),
    # reduced-size versions of BASELINE.json configs[3] and configs[4]
    "s4-foliage-small": (lambda S: S.foliage_field_scene(n_instances=100, n_blades=24, seg=6, n_point=180, n_quads=10, field=24.0), dict(spp_=8, res=(160, 90))),
    "s5-glass-knot-small-depth32": (lambda S: S.glass_knot_scene(nu=96, nv=24), dict(spp_=16, res=(96, 96))),
}


@pytest.mark.parametrize("name", list(CASES))
def test_image_matches_oracle(pkg, oracle, gpu_lib, name):
    make, kw = CASES[name]
    setup = make(pkg.scenes)
    img, stats, ref, ostats = _render_both(pkg, oracle, setup, **kw)
    assert np.isfinite(img).all()
    err = oracle.rel_mse(img, ref)
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"
    # same sampler streams => the ray budgets agree almost exactly (a handful of paths may
    # diverge through libm-level differences in shading)
    assert stats.camera_rays == ostats["camera_rays"]
    assert abs(int(stats.intersection_tests) - ostats["intersection_tests"]) <= 0.002 * ostats["intersection_tests"] + 8
    assert abs(int(stats.shadow_tests) - ostats["shadow_tests"]) <= 0.002 * ostats["shadow_tests"] + 8


def test_tile_and_sample_windows_sum_to_full_render(pkg, oracle, gpu_lib):
    """Splitting a render into tile ranges / sample ranges (the multi-GPU decomposition) and
    summing the {r,g,b,w} buffers reproduces the one-shot render."""
    setup = pkg.scenes.small_mixed_scene()
    integ = setup.make_integrator(spp_=8, res=(80, 48))
    sc = pkg.Scene(setup.flat)
    full, _ = sc.render(integ)
    nt = integ.n_tiles()
    acc = np.zeros_like(full)
    for tr in ((0, nt // 3), (nt // 3, nt)):
        for sr in ((0, 3), (3, 8)):
            sc.render(integ, rgbw=acc, tile_range=tr, sample_range=sr)
    sc.close()
    assert np.allclose(acc, full, rtol=2e-5, atol=2e-5)
    # weights are exact integers/filter values: independent of accumulation order up to fp32 sum order
    assert np.allclose(acc[:, 3], full[:, 3], rtol=1e-5)


def test_super_tile_numbering_windows(pkg, oracle, gpu_lib):
    """pbrt_b200_render_desc.tile_order: ranges of the super-tile numbering (what the multi-GPU work counter hands out) cover the
    frame exactly once -- partial edge tiles, positions past the image edge and the (0,2) sampler's per-tile seeds included."""
    setup = pkg.scenes.small_mixed_scene()
    for sampler_ in ("sobol", "02sequence"):
        integ = setup.make_integrator(spp_=4, res=(200, 120), sampler_=sampler_)  # 13 x 8 tiles: not a multiple of the super-tile edge
        sc = pkg.Scene(setup.flat)
        full, st = sc.render(integ)
        for S in (4, 8):
            n = integ.n_tile_positions(S)
            assert n == pkg.load_library().pbrt_b200_tile_positions(200, 120, S) and n > integ.n_tiles()
            acc = np.zeros_like(full)
            cams = 0
            for tr in ((0, n // 5), (n // 5, n // 2 + 3), (n // 2 + 3, n)):
                _, s2 = sc.render(integ, rgbw=acc, tile_range=tr, tile_order=S)
                cams += s2.camera_rays
            assert cams == st.camera_rays
            assert np.allclose(acc, full, rtol=2e-5, atol=2e-5)
        sc.close()


def test_small_paths_in_flight(pkg, oracle, gpu_lib):
    """Many small waves (capacity << work) give the same image as one big wave."""
    setup = pkg.scenes.cornell_scene()
    integ = setup.make_integrator(spp_=4, res=(64, 64))
    sc = pkg.Scene(setup.flat)
    a, _ = sc.render(integ, paths_in_flight=1024)
    b, _ = sc.render(integ)
    sc.close()
    assert np.allclose(a, b, rtol=2e-5, atol=2e-5)


# sphere area lights (Sphere::sample_interaction / pdf_wi): cone-sampled from outside, uniform-area from inside
# (the enclosing light), one-sided (emits only when hit); under the path integrator's three light strategies, the recursive
# integrators, and from a scene file
SPHERE_LIGHT_CASES = {
    "path-power": dict(spp_=16, res=(96, 64)),
    "path-spatial-halton": dict(spp_=8, res=(96, 64), strategy="spatial", sampler_="halton"),
    "path-uniform-02sequence": dict(spp_=4, res=(64, 48), strategy="uniform", sampler_="02sequence"),
}


@pytest.mark.parametrize("name", list(SPHERE_LIGHT_CASES))
def test_sphere_area_lights_match_oracle(pkg, oracle, gpu_lib, name):
    setup = pkg.scenes.sphere_lights_scene()
    img, stats, ref, ostats = _render_both(pkg, oracle, setup, **SPHERE_LIGHT_CASES[name])
    assert np.isfinite(img).all()
    err = oracle.rel_mse(img, ref)
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"
    assert stats.camera_rays == ostats["camera_rays"]
    assert abs(int(stats.shadow_tests) - ostats["shadow_tests"]) <= 0.002 * ostats["shadow_tests"] + 8


@pytest.mark.parametrize("kind", ["whitted", "all", "one"])
def test_sphere_area_lights_recursive_integrators(pkg, oracle, gpu_lib, kind):
    setup = pkg.scenes.sphere_lights_scene()
    base = setup.make_integrator(spp_=8, res=(80, 56))
    H = pkg.host
    integ = H.WhittedIntegrator(base.camera, base.film, base.sampler) if kind == "whitted" else H.DirectLightingIntegrator(base.camera, base.film, base.sampler, strategy=kind)
    sc = pkg.Scene(setup.flat)
    img, stats = integ.render(sc)
    sc.close()
    ref, ostats = oracle.render_image(setup.flat, integ)
    err = oracle.rel_mse(img, ref)
    assert err <= REL_MSE_TOL, f"relMSE {err:.3e}"
    assert stats.camera_rays == ostats["camera_rays"]


def test_sphere_area_light_from_a_scene_file(pkg, oracle, gpu_lib):
    text = '''LookAt 0 -6 2.5  0 0 .8  0 0 1
Camera "perspective" "float fov" 45
Film "image" "integer xresolution" [72] "integer yresolution" [48]
Sampler "sobol" "integer pixelsamples" 8
Integrator "path" "integer maxdepth" 4
WorldBegin
Material "matte" "rgb Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-8 -8 0 8 -8 0 8 8 0 -8 8 0]
AttributeBegin
  Translate 0 0 2.2
  AreaLightSource "diffuse" "rgb L" [9 8 7] "bool twosided" "true"
  Shape "sphere" "float radius" .4
AttributeEnd
AttributeBegin
  Translate -1.5 .5 .6
  Material "glass"
  Shape "sphere" "float radius" .6
AttributeEnd
WorldEnd
'''
    job = pkg.pbrt_parse_string(text).jobs[0]
    assert len(job.flat.lights) == 1 and job.flat.lights[0]["shape_kind"] == pkg.host.SHAPE_SPHERE and job.flat.lights[0]["two_sided"] == 1
    img, stats = job.render(device=0)
    ref, ostats = oracle.render_image(job.flat, job.integrator)
    assert oracle.rel_mse(img, ref) <= REL_MSE_TOL and stats.camera_rays == ostats["camera_rays"]
