"""CPU checks of the oracle's DirectLightingIntegrator / WhittedIntegrator restatement.  The reference has no
tests for these integrators, so the restatement is cross-checked through relations the algorithms guarantee:

* with delta lights only and flat-shaded surfaces, whitted.rs and directlighting.rs evaluate the very same expression per
  light (no MIS, one sample per light, same sampler dimensions) -> identical images, bit for bit;
* directlighting "one" is an unbiased one-light estimate of "all" -> they agree in expectation;
* maxdepth 1 turns the specular recursion off -> mirror / glass surfaces show direct lighting only (black for specular BSDFs);
* the sample-array bookkeeping of "all" (5 + 2 * maxdepth * n_lights * 2 array dimensions) moves where get_1d / get_2d read.
"""
import importlib

import numpy as np

pkg = importlib.import_module("pbrt-rust_b200")
H = pkg.host


def _integ(setup, kind, spp, res, sampler="sobol", maxdepth=5):
    base = setup.make_integrator(spp_=spp, res=res, sampler_=sampler)
    if kind == "whitted":
        return H.WhittedIntegrator(base.camera, base.film, base.sampler, maxdepth=maxdepth)
    return H.DirectLightingIntegrator(base.camera, base.film, base.sampler, maxdepth=maxdepth, strategy=kind)


def test_whitted_equals_directlighting_for_delta_lights(oracle):
    setup = pkg.scenes.spheres_scene()  # one distant light; matte quad, mirror and glass spheres (flat shading)
    a, sa = oracle.render_image(setup.flat, _integ(setup, "whitted", 4, (72, 72)))
    b, sb = oracle.render_image(setup.flat, _integ(setup, "all", 4, (72, 72)))
    # "all" reads its light samples from array dimensions, whitted from get_2d: a distant light ignores the sample, so the
    # images coincide exactly although the dimensions differ
    assert a.tobytes() == b.tobytes()
    assert sa["intersection_tests"] == sb["intersection_tests"] and sa["shadow_tests"] == sb["shadow_tests"]
    assert a.mean() > 0.05


def test_one_light_estimate_agrees_with_all_lights(oracle):
    setup = pkg.scenes.small_mixed_scene()
    one, _ = oracle.render_image(setup.flat, _integ(setup, "one", 64, (48, 32)))
    all_, _ = oracle.render_image(setup.flat, _integ(setup, "all", 64, (48, 32)))
    assert abs(one.mean() - all_.mean()) < 0.01 * all_.mean()
    assert oracle.rel_mse(one, all_) < 5e-3


def test_maxdepth_one_disables_the_specular_recursion(oracle):
    setup = pkg.scenes.spheres_scene()
    d1, s1 = oracle.render_image(setup.flat, _integ(setup, "whitted", 1, (64, 64), maxdepth=1))
    d5, s5 = oracle.render_image(setup.flat, _integ(setup, "whitted", 1, (64, 64), maxdepth=5))
    assert s1["intersection_tests"] == s1["camera_rays"] and s5["intersection_tests"] > s5["camera_rays"]
    sphere_px = (d1.sum(axis=2) == 0) & (d5.sum(axis=2) > 0)  # specular surfaces: black without recursion
    assert sphere_px.sum() > 200


def test_recursive_integrators_golden(oracle):
    """Frozen oracle outputs (tests/golden/recursive_golden.npz, made by tests/golden/make_recursive_golden.py)."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "recursive_golden.npz")
    setup = pkg.scenes.small_mixed_scene()
    for kind in ("whitted", "one", "all"):
        img, _ = oracle.render_image(setup.flat, _integ(setup, kind, 4, (48, 32)))
        assert np.allclose(img, g[kind], rtol=1e-5, atol=1e-6), kind
    img, _ = oracle.render_image(setup.flat, _integ(setup, "all", 4, (48, 32), sampler="halton", maxdepth=3))
    assert np.allclose(img, g["all_halton_d3"], rtol=1e-5, atol=1e-6)
    img, _ = oracle.render_image(setup.flat, _integ(setup, "all", 4, (48, 32), sampler="02sequence", maxdepth=3))
    assert np.allclose(img, g["all_zt_d3"], rtol=1e-5, atol=1e-6)


def test_02sequence_sample_arrays_shift_the_tile_stream(oracle):
    """Requesting 2D arrays makes start_pixel draw more from the tile's PCG32 (zerotwosequence.rs:67-71): with "all" the camera
    samples of every pixel after the first differ from the array-free "one" run; the first pixel's first sample does not."""
    setup = pkg.scenes.small_mixed_scene()
    one = _integ(setup, "one", 16, (32, 32), sampler="02sequence")
    all_ = _integ(setup, "all", 16, (32, 32), sampler="02sequence")
    a, _ = oracle.render_image(setup.flat, one)
    b, _ = oracle.render_image(setup.flat, all_)
    assert not np.allclose(a, b)
    assert abs(a.mean() - b.mean()) < 0.1 * a.mean()
