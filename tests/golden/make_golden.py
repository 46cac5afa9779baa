#!/usr/bin/env python3
"""Generates tests/golden/*.npz -- frozen outputs of the CPU oracle (oracle/) on seeded inputs.

The reference itself (nightly Rust + LALRPOP + ~40 crates) cannot be built in this image, so these vectors
are produced by the oracle AFTER it passed the reference's own tests (tests/test_oracle_reference_tests.py);
they pin the oracle against drift and give the GPU tests a fixture that does not need the oracle's source.
Regenerate with:  python tests/golden/make_golden.py   (deterministic; commit the .npz files it writes).
"""
import importlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def build():
    pkg = importlib.import_module("pbrt-rust_b200")
    from oracle import oracle as O
    S = pkg.scenes
    out = {}
    # 1. closest-hit / any-hit records on two scenes (gate 1 fixtures)
    for name, setup in (("mixed", S.small_mixed_scene()), ("cornell", S.cornell_scene())):
        flat = setup.flat
        rays = S.rays_diffuse(flat, 4096, seed=7)
        hits, cnt = O.intersect(flat, rays, nthreads=1)
        shadow = S.rays_shadow(flat, 4096, seed=11)
        occ, _ = O.intersect_p(flat, shadow, nthreads=1)
        out[f"{name}_hits"] = hits
        out[f"{name}_occluded"] = np.packbits(occ)
        out[f"{name}_counters"] = cnt
    # 2. sampler streams: Sobol, Halton, (0,2) -- camera sample + 3 x (1D, 2D) for a few pixels
    setup = S.small_mixed_scene()
    for samp in ("sobol", "halton", "02sequence"):
        integ = setup.make_integrator(spp_=8, res=(96, 64), sampler_=samp)
        streams = [O.sampler_stream(integ, seed=3, px=px, py=py, nsamples=8, n1d2d=3) for px, py in ((0, 0), (17, 5), (95, 63))]
        out[f"stream_{samp}"] = np.stack(streams)
    # 3. BSDF f / pdf / sample_f for the five materials (canonical frame)
    rows = []
    mats = [("matte", dict(Kd=(0.6, 0.3, 0.2))), ("matte", dict(Kd=0.5, sigma=20.0)), ("plastic", dict(Kd=(0.2, 0.3, 0.6), Ks=0.3, roughness=0.05)),
            ("mirror", dict(Kr=0.8)), ("glass", dict(index=1.5)), ("glass", dict(uroughness=0.1, vroughness=0.2)), ("metal", dict(roughness=0.05))]
    rng = np.random.RandomState(5)
    for m, kw in mats:
        row = pkg.host.SceneBuilder._mat_row(m, **kw)
        for _ in range(16):
            wo = rng.normal(size=3).astype(np.float32); wo /= np.linalg.norm(wo)
            wi = rng.normal(size=3).astype(np.float32); wi /= np.linalg.norm(wi)
            u = rng.uniform(size=2).astype(np.float32)
            rows.append(np.concatenate([wo, wi, u, O.bsdf_eval(row, wo, wi, u)]))
    out["bsdf"] = np.array(rows, np.float32)
    # 4. small images (gate 2 fixtures): linear RGB
    for name, setup, kw in (("mixed", S.small_mixed_scene(), dict(spp_=4, res=(48, 32))), ("spheres", S.spheres_scene(), dict(spp_=4, res=(40, 40))),
                            ("cornell", S.cornell_scene(), dict(spp_=4, res=(32, 32)))):
        img, st = O.render_image(setup.flat, setup.make_integrator(**kw), nthreads=1)
        out[f"image_{name}"] = img.astype(np.float32)
        out[f"image_{name}_rays"] = np.array([st["camera_rays"], st["intersection_tests"], st["shadow_tests"]], np.uint64)
    return out


if __name__ == "__main__":
    data = build()
    np.savez_compressed(HERE / "oracle_golden.npz", **data)
    print("wrote", HERE / "oracle_golden.npz", {k: v.shape for k, v in data.items()})
