"""The oracle's VolPathIntegrator restatement (oracle/oracle_volpath.hpp) pinned without the reference binary.

The reference holds no test for volpath.rs or homogeneous.rs and cannot be built here, so the restatement
is tied down from three sides:

  * to the already pinned PathIntegrator restatement: on a scene without media, without material-less surfaces and without
    specular lobes, volpath.rs:82-221 performs exactly the operations of path.rs:79-222 on exactly the same sample stream,
    so the two oracle functions must produce bit-identical images;
  * to closed forms of the transport they implement: Beer-Lambert attenuation of an emitter seen through an absorbing
    medium (HomogeneousMedium::sample's weight has expectation Tr), and the white furnace (a closed emitting box filled
    with a non-absorbing medium shows radiance 1 in every direction, whatever g: phase sampling, its pdf, the shadow-ray
    transmittance and the MIS weights have to agree for that);
  * to the reference's own deviation from pbrt-v3 (volpath.rs:131-135 `bounces -= 1; continue`): a medium behind a
    material-less boundary that the camera ray reaches at bounce 0 contributes nothing (the usize wraps, the path ends).
"""
import importlib
import math

import numpy as np
import pytest


def _volpath_of(pkg, integ, **kw):
    H = pkg.host
    return H.VolPathIntegrator(integ.camera, integ.film, integ.sampler, maxdepth=integ.max_depth, rrthreshold=integ.rr_threshold,
                               lightsamplestrategy=integ.light_sample_strategy, **kw)


def _lit_corner(pkg, sampler, emitter_kd):
    """Matte floor + wall, a plastic box, a quad area light whose own material is matte `emitter_kd`, a point light."""
    H, S = pkg.host, pkg.scenes
    b = H.SceneBuilder()
    for col, pts in (((0.6, 0.6, 0.6), ((-3, -3, 0), (3, -3, 0), (3, 3, 0), (-3, 3, 0))), ((0.2, 0.5, 0.7), ((-3, 3, 0), (3, 3, 0), (3, 3, 4), (-3, 3, 4)))):
        b.material("matte", Kd=col)
        P, I = S.quad(*pts)
        b.shape("trianglemesh", P=P, indices=I)
    b.material("plastic", Kd=(0.4, 0.3, 0.2), Ks=0.3, roughness=0.2)
    P, I = S.box_mesh((-0.8, -0.5, 0.0), (0.4, 0.7, 1.1))
    b.shape("trianglemesh", P=P, indices=I)
    b.attribute_begin()
    b.area_light_source("diffuse", L=(9, 8, 6))
    b.material("matte", Kd=emitter_kd)
    P, I = S.quad((-1, -1, 3.5), (-1, 1, 3.5), (1, 1, 3.5), (1, -1, 3.5))
    b.shape("trianglemesh", P=P, indices=I)
    b.attribute_end()
    b.light_source("point", **{"from": (2.0, -2.0, 2.0), "I": (4.0, 4.0, 4.0)})
    flat = b.world_end()
    film = H.Film(40, 40, "gaussian")
    cam = H.PerspectiveCamera(film, H.Transform.look_at((0.5, -5.0, 2.2), (0, 0, 1.2), (0, 0, 1)).inverse(), fov=50.0)
    return flat, H.PathIntegrator(cam, film, H.Sampler(sampler, 8), maxdepth=6, lightsamplestrategy="power")


@pytest.mark.parametrize("sampler,emitter_kd", [("sobol", 0.0), ("halton", 0.0), ("sobol", 0.3), ("02sequence", 0.3)])
def test_volpath_equals_path_without_media_or_specular_lobes(pkg, oracle, sampler, emitter_kd):
    # A BSDF WITHOUT a non-specular lobe (the black matte of an emitter: matte.rs adds no lobe for Kd = 0) is where the two differ:
    # path.rs:134 skips the light sample, volpath.rs:140 draws its five numbers and gets zero.  With the global samplers those are
    # trailing dimensions of a path that ends there (sample_f finds no lobe) and change nothing; the (0,2)-sequence sampler's draws
    # beyond its precomputed dimensions come from the tile's RNG and shift every later sample, so that case needs emitter_kd > 0.
    flat, integ = _lit_corner(pkg, sampler, emitter_kd)
    setup = type("S", (), {"flat": flat})
    a, sa = oracle.render(setup.flat, integ)
    b, sb = oracle.render(setup.flat, _volpath_of(pkg, integ))
    assert np.array_equal(a, b)
    assert sa["camera_rays"] == sb["camera_rays"]
    # the shadow rays of path.rs are intersect_p calls; volpath's transmittance loop uses intersect (light.rs:133)
    assert sb["shadow_tests"] == 0 and sb["intersection_tests"] == sa["intersection_tests"] + sa["shadow_tests"]


def _emitter_behind_slab(pkg, sigma_a, sigma_s, dist, L=(2.0, 3.0, 4.0), spp=64, res=8, g=0.0, maxdepth=5):
    """Camera at the origin inside medium `m`, looking down +z at a large emissive quad at z = dist; nothing else."""
    H = pkg.host
    b = H.SceneBuilder()
    b.make_named_medium("m", sigma_a=sigma_a, sigma_s=sigma_s, g=g)
    b.medium_interface("m", "m")  # no transitions: every ray stays in the medium
    b.area_light_source("diffuse", L=L, twosided=True)
    b.material("matte", Kd=0.0)
    P, I = pkg.scenes.quad((-50, -50, dist), (50, -50, dist), (50, 50, dist), (-50, 50, dist))
    b.shape("trianglemesh", P=P, indices=I)
    cam_medium = b.camera_medium()
    flat = b.world_end()
    film = H.Film(res, res, "box")
    cam = H.PerspectiveCamera(film, H.Transform.look_at((0, 0, 0), (0, 0, 1), (0, 1, 0)).inverse(), fov=2.0)
    return flat, H.VolPathIntegrator(cam, film, H.Sampler("sobol", spp), maxdepth=maxdepth, lightsamplestrategy="uniform", camera_medium=cam_medium)


def test_beer_lambert_through_a_grey_absorber(pkg, oracle):
    # grey sigma_t: every unscattered sample carries exactly Tr / pdf = 1 and survives with probability Tr
    sig, dist = 0.35, 2.0
    flat, integ = _emitter_behind_slab(pkg, (sig,) * 3, (0.0,) * 3, dist, spp=1024)
    img, _ = oracle.render_image(flat, integ)
    want = np.array([2.0, 3.0, 4.0]) * math.exp(-sig * dist)
    got = img.reshape(-1, 3).mean(0)
    # 8 x 8 x 1024 Bernoulli(Tr = 0.497) samples: sigma = sqrt(p (1 - p) / n) / p = 0.4 %
    assert np.allclose(got, want, rtol=0.02), (got, want)


def test_beer_lambert_through_a_chromatic_absorber(pkg, oracle):
    # chromatic sigma_t: the single-channel distance sampling is weighted by the channel-averaged pdf (homogeneous.rs:59-71);
    # its expectation is still the per-channel transmittance
    sig, dist = (0.1, 0.5, 1.2), 1.5
    flat, integ = _emitter_behind_slab(pkg, sig, (0.0,) * 3, dist, spp=2048)
    img, _ = oracle.render_image(flat, integ)
    want = np.array([2.0, 3.0, 4.0]) * np.exp(-np.array(sig) * dist)
    got = img.reshape(-1, 3).mean(0)
    assert np.allclose(got, want, rtol=0.03), (got, want)


def _furnace(pkg, sigma_s, g, spp=1024, maxdepth=90, res=6):
    """A closed cube whose six inner faces emit L = 1 (black matte, so nothing is reflected), filled with a non-absorbing medium."""
    H = pkg.host
    b = H.SceneBuilder()
    b.make_named_medium("m", sigma_a=(0.0,) * 3, sigma_s=sigma_s, g=g)
    b.medium_interface("m", "m")
    b.area_light_source("diffuse", L=(1.0, 1.0, 1.0), twosided=True)
    b.material("matte", Kd=0.0)
    P, I = pkg.scenes.box_mesh((-1, -1, -1), (1, 1, 1))
    b.shape("trianglemesh", P=P, indices=I)
    cam_medium = b.camera_medium()
    flat = b.world_end()
    film = H.Film(res, res, "box")
    cam = H.PerspectiveCamera(film, H.Transform.look_at((0.1, -0.2, 0.05), (0.7, 0.3, 1.0), (0, 1, 0)).inverse(), fov=70.0)
    return flat, H.VolPathIntegrator(cam, film, H.Sampler("halton", spp), maxdepth=maxdepth, rrthreshold=0.0, lightsamplestrategy="uniform",
                                     camera_medium=cam_medium)


@pytest.mark.parametrize("sigma_s,g", [((0.8, 0.8, 0.8), 0.0), ((0.5, 1.0, 1.5), 0.7), ((1.2, 1.2, 1.2), -0.5)])
def test_white_furnace_with_a_scattering_medium(pkg, oracle, sigma_s, g):
    flat, integ = _furnace(pkg, sigma_s, g)
    img, _ = oracle.render_image(flat, integ)
    got = img.reshape(-1, 3).mean(0)
    # radiance 1 whatever the medium does: emission seen directly at bounce 0, otherwise gathered by light sampling + MIS at every
    # scattering vertex; maxdepth 90 (900 of Halton's 1000 dimensions) truncates nothing visible at optical thickness ~2.  Measured with
    # 8 x 8 x 8192 samples: 1.0004 +- 0.002 (sigma_s 1.5, g 0); here the worst channel's standard error is ~0.02
    assert np.allclose(got, 1.0, atol=0.06), got


def test_medium_behind_a_boundary_hit_at_bounce_zero_is_invisible(pkg, oracle):
    # volpath.rs:131-135 (see oracle_volpath.hpp): the camera ray crosses a material-less surface at bounce 0, `bounces -= 1` wraps and
    # the path ends at its next depth test -- a lit, dense smoke cube in front of a black background renders exactly black
    H = pkg.host
    b = H.SceneBuilder()
    b.make_named_medium("smoke", sigma_a=(0.1,) * 3, sigma_s=(4.0,) * 3)
    b.light_source("point", **{"from": (0.0, 0.0, 3.0), "I": (50.0, 50.0, 50.0)})
    b.attribute_begin()
    b.medium_interface("smoke", "")
    b.material("none")
    P, I = pkg.scenes.box_mesh((-1, -1, -1), (1, 1, 1))
    b.shape("trianglemesh", P=P, indices=I)
    b.attribute_end()
    flat = b.world_end()
    film = H.Film(16, 16, "box")
    cam = H.PerspectiveCamera(film, H.Transform.look_at((0, -5, 0.5), (0, 0, 0), (0, 0, 1)).inverse(), fov=30.0)
    integ = H.VolPathIntegrator(cam, film, H.Sampler("sobol", 16), maxdepth=5, lightsamplestrategy="uniform")
    rgbw, st = oracle.render(flat, integ)
    assert st["camera_rays"] == 16 * 16 * 16 and np.all(rgbw[:, :3] == 0.0)
    # the same cube seen from INSIDE a surrounding medium is reached after real bounces and does light up
    setup = pkg.scenes.fog_box_scene(xres=32, yres=32, spp=16)
    img, _ = oracle.render_image(setup.flat, setup.make_integrator())
    assert img.mean() > 0.1 and np.isfinite(img).all()


def test_medium_rows_and_interfaces_of_the_fog_box(pkg):
    # flattening: media in MakeNamedMedium order (sigma * scale), interfaces per primitive row, the camera medium
    H = pkg.host
    setup = pkg.scenes.fog_box_scene()
    flat = setup.flat
    assert len(flat.media) == 3 and np.allclose(flat.media[1]["sigma_s"], np.array([2.5, 2.6, 2.8], np.float32) * np.float32(1.5)) and flat.media[1]["g"] == np.float32(0.6)
    pm = flat.prim_media[np.argsort(flat.prims["creation_index"])]
    assert len(pm) == len(flat.prims)
    assert tuple(pm[0]) == (-1, 0) and tuple(pm[1]) == (-1, 0)       # floor: outermost state ("", haze) -- a transition
    assert tuple(pm[2]) == (0, 0)                                       # walls: (haze, haze) -- not a transition
    smoke = [tuple(r) for r, p in zip(pm, flat.prims[np.argsort(flat.prims["creation_index"])]) if p["material"] < 0]
    assert len(smoke) == 12 and set(smoke) == {(1, 0)}
    sph = [tuple(r) for r, p in zip(pm, flat.prims[np.argsort(flat.prims["creation_index"])]) if p["shape_kind"] == H.SHAPE_SPHERE]
    assert sph == [(2, 0)]
    assert setup.make_integrator().camera_medium == 0
    inst = pkg.scenes.fog_box_scene(instanced=True).flat
    assert len(inst.prim_media) == len(inst.prims) and (inst.prim_media["inside"] == 1).sum() == 12
    assert tuple(inst.prim_media[inst.prims["shape_kind"] == H.SHAPE_INSTANCE][0]) == (-1, -1)


def test_volpath_scene_file_round_trip(pkg, oracle, tmp_path):
    # MakeNamedMedium / MediumInterface / Integrator "volpath" through the scene-file front end (api.rs:1211-1258): the written file
    # parses back to the same tables, camera medium included (Camera.medium = outside medium of the graphics state at WorldEnd)
    setup = pkg.scenes.fog_box_scene(xres=24, yres=24, spp=4)
    integ = setup.make_integrator()
    path = tmp_path / "fog.pbrt"
    importlib.import_module("pbrt-rust_b200.scenefile").write_pbrt(path, setup.flat, integ)
    job = pkg.pbrt_parse(path).jobs[0]
    assert job.integrator.name == "volpath" and job.integrator.camera_medium == integ.camera_medium
    assert np.array_equal(job.flat.media, setup.flat.media) and np.array_equal(job.flat.prim_media, setup.flat.prim_media)
    a, _ = oracle.render(setup.flat, integ)
    b, _ = oracle.render(job.flat, job.integrator)
    assert np.array_equal(a, b)
