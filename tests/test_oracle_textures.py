"""Textures, MIPMaps and camera ray differentials of the CPU oracle (oracle/oracle_texture.hpp) and of the host-side pyramid builder
(pbrt-rust_b200/textures.py) against closed forms and an INDEPENDENT numpy restatement written from the formulas of the reference
(src/core/texture.rs, src/textures/*.rs, src/core/mipmap.rs) -- the reference holds no tests for textures, so this is
what pins the oracle's texture half; the GPU then matches the oracle (tests/test_gpu_textures.py).

The evaluator below walks the Tex TREE (not the flattened postfix program), so it also checks TextureTables' flattening."""
import importlib

import numpy as np
import pytest

f32 = np.float32


@pytest.fixture(scope="module")
def T():
    return importlib.import_module("pbrt-rust_b200.textures")


def _flat_with(pkg, T, tex):
    """A FlatScene holding just the program of `tex` -> (flat, texref)."""
    tabs = T.TextureTables()
    ref = tabs.program(tex)
    fs = pkg.host.FlatScene()
    fs.textures = tabs.node_array()
    fs.mipmap_objects = list(tabs.mipmaps)
    fs.mipmaps = tabs.mipmap_array()
    return fs, ref


# ---- the independent evaluator (float64, recursive over the tree) -------------------------------------------------------------
def _xf_point(m, p):
    m = m.reshape(4, 4).astype(np.float64)
    q = p @ m[:3, :3].T + m[:3, 3]
    w = p @ m[3, :3] + m[3, 3]
    return q / w[:, None]


def _map2d(T, mp, p, uv):
    k, m = mp.kind, mp.m.astype(np.float64)
    if k == T.MAP_UV:
        return np.stack([m[0] * uv[:, 0] + m[2], m[1] * uv[:, 1] + m[3]], axis=1)
    if k == T.MAP_PLANAR:
        return np.stack([m[6] + p @ m[0:3], m[7] + p @ m[3:6]], axis=1)
    v = _xf_point(mp.m, p)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if k == T.MAP_SPHERICAL:
        phi = np.arctan2(v[:, 1], v[:, 0])
        phi = np.where(phi < 0, phi + 2 * np.pi, phi)
        return np.stack([np.arccos(np.clip(v[:, 2], -1, 1)) / np.pi, phi / (2 * np.pi)], axis=1)
    return np.stack([np.pi + np.arctan2(v[:, 1], v[:, 0]) / (2 * np.pi), v[:, 2]], axis=1)  # cylindrical, texture.rs:222-229


def _eval_tree(T, t, p, uv):
    n = len(p)
    if not isinstance(t, T.Tex):
        return np.broadcast_to(np.asarray(t, np.float64).reshape(-1)[[0, 0, 0]] if np.ndim(t) == 0 else np.asarray(t, np.float64), (n, 3)).copy()
    kids = [_eval_tree(T, c, p, uv) for c in t.children]
    if t.kind == T.TEX_CONSTANT:
        return np.broadcast_to(t.v[:3].astype(np.float64), (n, 3)).copy()
    if t.kind == T.TEX_SCALE:
        return kids[0] * kids[1]
    if t.kind == T.TEX_MIX:
        a = kids[2][:, :1]
        return kids[0] * (1 - a) + kids[1] * a
    if t.kind == T.TEX_BILERP:
        st = _map2d(T, t.mapping, p, uv)
        v = t.v.astype(np.float64).reshape(4, 3)
        s, tt = st[:, :1], st[:, 1:]
        return v[0] * (1 - tt) * (1 - s) + v[1] * (1 - s) * tt + v[2] * (1 - tt) * s + v[3] * tt * s
    if t.kind == T.TEX_UV:
        st = _map2d(T, t.mapping, p, uv)
        return np.stack([st[:, 0] - np.floor(st[:, 0]), st[:, 1] - np.floor(st[:, 1]), np.zeros(n)], axis=1)
    if t.kind == T.TEX_CHECKERBOARD2D:  # aamode none
        st = _map2d(T, t.mapping, p, uv)
        first = ((np.floor(st[:, 0]) + np.floor(st[:, 1])) % 2 == 0)[:, None]
        return np.where(first, kids[0], kids[1])
    if t.kind == T.TEX_CHECKERBOARD3D:
        q = _xf_point(t.m, p)
        first = (np.floor(q).sum(axis=1) % 2 == 0)[:, None]
        return np.where(first, kids[0], kids[1])
    raise NotImplementedError(t.kind)


def _points(n, seed, lo=-3.0, hi=3.0):
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, (n, 3)).astype(f32), rng.uniform(-2.0, 3.0, (n, 2)).astype(f32)


def _away_from_cell_edges(st, eps=1e-3):
    f = st - np.floor(st)
    return ((f > eps) & (f < 1 - eps)).all(axis=1)


def test_expression_trees_match_an_independent_evaluator(pkg, oracle, T):
    H = pkg.host
    M2 = T.Mapping2D
    w2t = (H.Transform.translate((0.3, -0.2, 0.5)) * H.Transform.rotate(33.0, (0.2, 1.0, 0.3)) * H.Transform.scale(0.7, 1.3, 0.9))
    maps = [M2.uv(3.0, 2.0, 0.25, -0.5), M2.planar((0.7, 0.1, 0.0), (0.0, 0.2, 0.9), 0.3, 0.1), M2.spherical(w2t), M2.cylindrical(w2t)]
    p, uv = _points(4000, 1)
    for k, mp in enumerate(maps):
        inner = T.Tex.bilerp(mp, (0.1, 0.2, 0.3), (0.9, 0.1, 0.4), (0.3, 0.8, 0.2), (0.6, 0.6, 0.9))
        tree = T.Tex.mix(T.Tex.scale(inner, np.array([0.5, 1.5, 2.0], f32)), T.Tex.uv(mp), T.Tex.checkerboard(mp, 0.2, 0.7, "none"))
        tree = T.Tex.checkerboard3d(w2t, tree, T.Tex.checkerboard(M2.uv(5.0, 4.0), np.array([1.0, 0.5, 0.25], f32), inner, "none"))
        flat, ref = _flat_with(pkg, T, tree)
        got = oracle.texture_eval(flat, ref, p, uv=uv)
        want = _eval_tree(T, tree, p.astype(np.float64), uv.astype(np.float64))
        # points within 1e-3 of a check boundary may legitimately land on the other side in f32
        ok = _away_from_cell_edges(_map2d(T, mp, p.astype(np.float64), uv.astype(np.float64))) & _away_from_cell_edges(_xf_point(w2t.m, p.astype(np.float64))) & \
            _away_from_cell_edges(np.stack([5.0 * uv[:, 0], 4.0 * uv[:, 1]], axis=1).astype(np.float64))
        assert ok.mean() > 0.9
        assert np.allclose(got[ok], want[ok], rtol=2e-4, atol=2e-5), (k, np.abs(got[ok] - want[ok]).max())


def test_postfix_programs_share_equal_trees_and_report_their_stack_depth(pkg, T):
    mp = T.Mapping2D.uv()
    a = T.Tex.mix(T.Tex.checkerboard(mp, 0.1, 0.9), T.Tex.uv(mp), T.Tex.scale(0.5, T.Tex.checkerboard(mp, 0.0, 1.0)))
    b = T.Tex.mix(T.Tex.checkerboard(mp, 0.1, 0.9), T.Tex.uv(mp), T.Tex.scale(0.5, T.Tex.checkerboard(mp, 0.0, 1.0)))
    tabs = T.TextureTables()
    assert tabs.program(a) == tabs.program(b) and len(tabs.nodes) == 10
    kinds = [int(n["kind"]) for n in tabs.nodes]
    assert kinds[-1] == T.TEX_MIX and kinds.count(T.TEX_CONSTANT) == 5
    assert tabs.max_depth == 5  # checker, uv, 0.5, and the two operands of the inner checkerboard
    deep = T.Tex.constant(1.0)
    for _ in range(9):
        deep = T.Tex.scale(0.5, deep)
    b2 = pkg.host.SceneBuilder()
    b2.material("matte", Kd=deep)
    b2.shape("sphere", radius=1.0)
    with pytest.raises(pkg.B200Error, match="value stack"):
        b2.world_end()


def test_closed_form_checkerboard_filter(pkg, oracle, T):
    mp = T.Mapping2D.uv(1.0, 1.0)
    tex = T.Tex.checkerboard(mp, 1.0, 0.0, "closedform")
    flat, ref = _flat_with(pkg, T, tex)
    uv = np.array([[0.5, 0.5], [1.5, 0.5], [0.5, 0.5], [0.98, 0.5], [0.5, 0.5]], f32)
    duv = np.array([[0.1, 0, 0, 0.1],     # footprint inside one check: point sample of check (0,0) -> tex1
                    [0.1, 0, 0, 0.1],     # check (1,0) -> tex2
                    [3.0, 0, 0, 0.1],     # wider than a check in s: the 50 % average (checkerboard.rs:64)
                    [0.1, 0, 0, 0.1],     # straddles the s = 1 edge: the reference's area2 = sint*tint - 2*sint*tint expression
                    [0.0, 0, 0, 0.0]], f32)
    got = oracle.texture_eval(flat, ref, np.zeros((5, 3), f32), uv=uv, duv=duv)[:, 0]
    assert got[0] == 1.0 and got[1] == 0.0 and got[2] == 0.5 and got[4] == 1.0

    def bump(x):
        return np.floor(x / 2) + 2 * np.maximum(x / 2 - np.floor(x / 2) - 0.5, 0)

    sint = (bump(0.98 + 0.1) - bump(0.98 - 0.1)) / 0.2
    tint = (bump(0.6) - bump(0.4)) / 0.2
    area2 = sint * tint - 2 * sint * tint
    assert abs(got[3] - (1.0 - area2)) < 1e-5


# ---- noise ------------------------------------------------------------------------------------------------------------------
def _perm():
    txt = open(importlib.import_module("pbrt-rust_b200").host._HERE / "csrc" / "texture.cuh").read()
    body = txt[txt.index("c_noise_perm[512] = {") + len("c_noise_perm[512] = {"):]
    body = body[: body.index("};")]
    return np.array([int(v) for v in body.replace("\n", " ").split(",")], np.int64)


def _noise_np(perm, x, y, z):
    """Perlin noise as in texture.rs:330-383 for NON-NEGATIVE coordinates, float64."""
    ix, iy, iz = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64), np.floor(z).astype(np.int64)
    dx, dy, dz = x - ix, y - iy, z - iz
    ix &= 255; iy &= 255; iz &= 255

    def grad(a, b, c, u_, v_, w_):
        h = perm[perm[perm[a] + b] + c] & 15
        u = np.where((h < 8) | (h == 12) | (h == 13), u_, v_)
        v = np.where((h < 4) | (h == 12) | (h == 13), v_, w_)
        return np.where(h & 1, -u, u) + np.where(h & 2, -v, v)

    def wgt(t):
        return 6 * t ** 5 - 15 * t ** 4 + 10 * t ** 3

    def lerp(t, a, b):
        return a * (1 - t) + b * t

    w = {}
    for i in (0, 1):
        for j in (0, 1):
            for k in (0, 1):
                w[i, j, k] = grad(ix + i, iy + j, iz + k, dx - i, dy - j, dz - k)
    wx, wy, wz = wgt(dx), wgt(dy), wgt(dz)
    x00, x10, x01, x11 = lerp(wx, w[0, 0, 0], w[1, 0, 0]), lerp(wx, w[0, 1, 0], w[1, 1, 0]), lerp(wx, w[0, 0, 1], w[1, 0, 1]), lerp(wx, w[0, 1, 1], w[1, 1, 1])
    return lerp(wz, lerp(wy, x00, x10), lerp(wy, x01, x11))


def test_noise_based_textures(pkg, oracle, T):
    H = pkg.host
    perm = _perm()
    assert len(perm) == 512 and np.array_equal(perm[:256], perm[256:]) and sorted(perm[:256]) == list(range(256))  # a doubled permutation of 0..255
    ident = H.Transform()
    # fbm with zero differentials runs all its octaves (log2(0) = -inf clamps to max_octaves): sum of omega^i noise(1.99^i p) + nothing more
    octaves, omega = 5, 0.6
    flat, ref = _flat_with(pkg, T, T.Tex.fbm(ident, octaves, omega))
    rng = np.random.default_rng(3)
    p = rng.uniform(0.5, 40.0, (3000, 3)).astype(f32)
    got = oracle.texture_eval(flat, ref, p)[:, 0].astype(np.float64)
    want = np.zeros(len(p))
    lam, o = 1.0, 1.0
    for _ in range(octaves):
        q = p.astype(np.float64) * lam
        want += o * _noise_np(perm, q[:, 0], q[:, 1], q[:, 2])
        lam *= 1.99; o *= omega
    # (the partial octave after the loop has weight smooth_step(0.3, 0.7, 0) = 0)
    assert np.allclose(got, want, atol=2e-4), np.abs(got - want).max()
    assert np.abs(got).max() < 2.5 and got.std() > 0.1
    # the noise function vanishes on the integer lattice
    lattice = np.array([[1, 2, 3], [17, 5, 250], [255, 255, 255]], f32)
    flat1, ref1 = _flat_with(pkg, T, T.Tex.fbm(ident, 1, 1.0))
    assert np.allclose(oracle.texture_eval(flat1, ref1, lattice), 0.0, atol=1e-6)
    # wide differentials switch octaves off: |dpdx| = 1 -> n = clamp(-1 - 0.5 * log2(1), 0, ..) = 0 octaves -> 0
    wide = oracle.texture_eval(flat, ref, p[:10], dpdx=np.tile([1.0, 0, 0], (10, 1)), dpdy=np.tile([0, 1.0, 0], (10, 1)))
    assert np.allclose(wide, 0.0)
    # REFERENCE QUIRK, reproduced on purpose: `x.floor() as usize` (texture.rs:332-334) saturates negative cells to 0, so the offset
    # dx = x - 0 is negative and the quintic weights explode -- noise is only usable in the positive octant
    neg = oracle.texture_eval(flat1, ref1, np.array([[-3.3, 0.5, 0.5]], f32))[0, 0]
    assert abs(neg) > 100.0
    # windy = |fbm(0.1 p, 0.5, 3)| * fbm(p, 0.5, 6); wrinkled (turbulence) adds o + |noise| per octave as the reference writes it
    flatw, refw = _flat_with(pkg, T, T.Tex.windy(ident))
    gw = oracle.texture_eval(flatw, refw, p[:500])[:, 0].astype(np.float64)

    def fbm_np(q, om, n):
        s, lam, o = np.zeros(len(q)), 1.0, 1.0
        for _ in range(n):
            s += o * _noise_np(perm, q[:, 0] * lam, q[:, 1] * lam, q[:, 2] * lam); lam *= 1.99; o *= om
        return s

    q = p[:500].astype(np.float64)
    assert np.allclose(gw, np.abs(fbm_np(q * 0.1, 0.5, 3)) * fbm_np(q, 0.5, 6), atol=2e-4)
    flatr, refr = _flat_with(pkg, T, T.Tex.wrinkled(ident, 3, 0.5))
    gr = oracle.texture_eval(flatr, refr, p[:500])[:, 0].astype(np.float64)
    want = np.zeros(500); lam, o = 1.0, 1.0
    for _ in range(3):
        want += o + np.abs(_noise_np(perm, q[:, 0] * lam, q[:, 1] * lam, q[:, 2] * lam)); lam *= 1.99; o *= 0.5
    want += o + 0.2  # lerp(smooth_step(.3, .7, 0) = 0, 0.2, |noise|) = 0.2 ; no octaves left after nint = max
    assert np.allclose(gr, want, atol=2e-4)
    # dots: inside / outside are swapped by create_dots_* (dots.rs:59-70); the struct's semantics are what the node carries
    flatd, refd = _flat_with(pkg, T, T.Tex.dots(T.Mapping2D.uv(), outside=0.25, inside=0.75))
    uv = rng.uniform(0.0, 12.0, (4000, 2)).astype(f32)
    gd = oracle.texture_eval(flatd, refd, np.zeros((4000, 3), f32), uv=uv)[:, 0]
    assert set(np.unique(gd)) == {f32(0.25), f32(0.75)} and 0.05 < (gd == f32(0.75)).mean() < 0.35  # pi r^2 = 0.385 of half the cells


# ---- MIPMap --------------------------------------------------------------------------------------------------------------------
def test_pyramid_of_a_power_of_two_image_is_box_filtered(T):
    rng = np.random.default_rng(5)
    img = rng.random((16, 32, 3)).astype(f32)
    for wrap in ("repeat", "clamp", "black"):
        m = T.MipMap(img, wrap=wrap)
        assert (m.width, m.height, m.n_levels) == (32, 16, 6) and [l.shape[:2] for l in m.levels] == [(16, 32), (8, 16), (4, 8), (2, 4), (1, 2), (1, 1)]
        assert np.array_equal(m.levels[0], img)
        l1 = img.reshape(8, 2, 16, 2, 3).astype(np.float64).mean(axis=(1, 3))
        assert np.allclose(m.levels[1], l1, rtol=1e-6)
        # levels whose height is already 1 read texel (s, 1) outside the image: wrap decides (mipmap.rs:301-321)
        prev = m.levels[4].astype(np.float64)  # 1 x 2
        if wrap == "black":
            want = (prev[0, 0] + prev[0, 1]) * 0.25
        else:
            want = (prev[0, 0] + prev[0, 1]) * 0.5
        assert np.allclose(m.levels[5][0, 0], want, rtol=1e-6)
        assert len(m.texels) == sum(l.size for l in m.levels)


def test_resampling_to_a_power_of_two(T):
    # Lanczos weights are normalised: a constant image stays constant (up to rounding), whatever the wrap mode does at the border
    const = np.full((5, 12, 3), 0.37, f32)
    m = T.MipMap(const, wrap="repeat")
    assert (m.width, m.height) == (16, 8) and np.allclose(m.levels[0], 0.37, rtol=1e-5) and np.allclose(m.levels[-1], 0.37, rtol=1e-5)
    first, w = T.resample_weights(12, 16)
    assert np.allclose(w.sum(axis=1), 1.0, atol=1e-6) and first[0] == -2 and first[-1] == 10  # floor(center - 2 + 0.5), center = (i + .5) * 12 / 16
    # a smooth image survives the resampling: a horizontal ramp stays a ramp away from the borders
    ramp = np.tile(np.linspace(0.0, 1.0, 24, dtype=f32)[None, :, None], (4, 1, 1))
    r = T.MipMap(ramp, wrap="clamp").levels[0][0, :, 0]
    x = (np.arange(32) + 0.5) * 24 / 32 - 0.5
    assert np.allclose(r[4:-4], (x / 23.0)[4:-4], atol=0.01)
    assert (T.MipMap(np.full((3, 3, 1), -1.0, f32)).levels[0] >= 0).all()  # the t pass clamps at zero (mipmap.rs:136)


def test_image_map_lookups(pkg, oracle, T):
    rng = np.random.default_rng(9)
    img = rng.random((16, 16, 3)).astype(f32)
    uvmap = T.Mapping2D.uv()
    centres = (np.stack(np.meshgrid(np.arange(16), np.arange(16)), axis=-1).reshape(-1, 2) + 0.5) / 16.0  # (s, t) of every texel centre
    zeros = np.zeros((len(centres), 3), f32)
    for trilinear in (False, True):
        for wrap in ("repeat", "clamp", "black"):
            mip = T.MipMap(img, do_trilinear=trilinear, wrap=wrap)
            flat, ref = _flat_with(pkg, T, T.Tex.imagemap(uvmap, mip))
            # zero differentials: bilinear lookup of level 0 (mipmap.rs:213, :248), exact at texel centres
            got = oracle.texture_eval(flat, ref, zeros, uv=centres)
            assert np.allclose(got.reshape(16, 16, 3), img, rtol=1e-5, atol=1e-6), (trilinear, wrap)
            # halfway between two texel centres: their average
            mid = oracle.texture_eval(flat, ref, zeros[:1], uv=[[4.0 / 16.0, 4.5 / 16.0]])[0]
            assert np.allclose(mid, 0.5 * (img[4, 3].astype(np.float64) + img[4, 4]), rtol=1e-5)
            # a footprint as wide as the image: the coarsest level, i.e. the image mean (power-of-two box filtering)
            wide = oracle.texture_eval(flat, ref, zeros[:1], uv=[[0.3, 0.6]], duv=[[1.0, 0.0, 0.0, 1.0]])[0]
            if trilinear or wrap != "black":  # (EWA over the 1 x 1 level reaches outside the image, where Black wrap returns zeros)
                assert np.allclose(wide, img.astype(np.float64).mean(axis=(0, 1)), rtol=1e-4), (trilinear, wrap)
            else:
                assert (wide < img.mean(axis=(0, 1))).all() and (wide > 0).all()
    # trilinear between two levels: width 2 / 16 -> level 1 exactly; width between -> a blend of the two bilinear lookups
    mip = T.MipMap(img, do_trilinear=True)
    flat, ref = _flat_with(pkg, T, T.Tex.imagemap(uvmap, mip))
    st = np.array([[5.0 / 16.0, 7.0 / 16.0]], f32)  # a corner shared by four texels of level 0 = centre of nothing; in level 1: texel centre
    l1 = oracle.texture_eval(flat, ref, zeros[:1], uv=[[(2 * 2 + 1) / 16.0, (2 * 3 + 1) / 16.0]], duv=[[2.0 / 16.0, 0, 0, 0]])[0]
    assert np.allclose(l1, mip.levels[1][3, 2], rtol=1e-5)
    # EWA: an isotropic footprint of one texel at a texel centre weights that texel most and stays within the local range
    mipe = T.MipMap(img, do_trilinear=False)
    flate, refe = _flat_with(pkg, T, T.Tex.imagemap(uvmap, mipe))
    e = oracle.texture_eval(flate, refe, zeros[:1], uv=[[8.5 / 16.0, 8.5 / 16.0]], duv=[[1.0 / 16.0, 0, 0, 1.0 / 16.0]])[0]
    nb = img[6:11, 6:11].reshape(-1, 3)
    assert (e >= nb.min(axis=0) - 1e-6).all() and (e <= nb.max(axis=0) + 1e-6).all()
    # ... and a constant image is a constant under every filter and footprint
    mc = T.MipMap(np.full((8, 8, 1), 0.6, f32), do_trilinear=False, max_anisotropy=4.0)
    flatc, refc = _flat_with(pkg, T, T.Tex.imagemap(uvmap, mc))
    duv = rng.uniform(-0.5, 0.5, (200, 4)).astype(f32)
    c = oracle.texture_eval(flatc, refc, np.zeros((200, 3), f32), uv=rng.random((200, 2)).astype(f32), duv=duv)
    assert np.allclose(c, 0.6, rtol=1e-5)
    # float image maps carry their value in all three channels; black wrap returns zero outside [0, 1)
    mb = T.MipMap(np.full((4, 4, 1), 0.5, f32), wrap="black")
    flatb, refb = _flat_with(pkg, T, T.Tex.imagemap(uvmap, mb))
    out = oracle.texture_eval(flatb, refb, np.zeros((2, 3), f32), uv=[[0.5, 0.5], [1.7, 0.5]])
    assert np.allclose(out[0], 0.5) and np.allclose(out[1], 0.0)


def test_texel_conversion_and_image_files(T, tmp_path):
    from PIL import Image

    rgb8 = (np.arange(4 * 6 * 3).reshape(4, 6, 3) * 3 % 256).astype(np.uint8)
    Image.fromarray(rgb8).save(tmp_path / "a.png")
    img = T.read_image(tmp_path / "a.png")
    assert img.shape == (4, 6, 3) and np.array_equal(img, (rgb8.astype(f32) / f32(255.0)))
    tex = T.convert_texels(img, as_float=False, scale=2.0, gamma=True)
    assert np.allclose(tex[0], 2.0 * T.inverse_gamma_correct(img[-1]))  # flipped in y (imagemap.rs:147-153)
    lo = f32(0.03) / f32(12.92)
    assert np.isclose(T.inverse_gamma_correct(np.array([0.03], f32))[0], lo) and np.isclose(T.inverse_gamma_correct(np.array([1.0], f32))[0], 1.0)
    y = T.convert_texels(img, as_float=True, scale=1.0, gamma=False)
    assert y.shape == (4, 6, 1) and np.allclose(y[::-1, :, 0], img @ np.array([0.212671, 0.715160, 0.072169], f32), rtol=1e-6)
    # PFM: bottom-up rows, little- or big-endian by the sign of the scale
    data = np.arange(2 * 3 * 3, dtype=f32).reshape(2, 3, 3)
    for endian, scale in (("<", "-1.0"), (">", "2.0")):
        (tmp_path / "b.pfm").write_bytes(b"PF\n3 2\n" + scale.encode() + b"\n" + data[::-1].astype(endian + "f4").tobytes())
        got = T.read_image(tmp_path / "b.pfm")
        assert np.array_equal(got, data * abs(float(scale)))
    # the cache is keyed like TexInfo (imagemap.rs:36-62)
    T.clear_cache()
    a = T.image_mipmap(str(tmp_path / "a.png"), False, False, 8.0, "repeat", 1.0, True)
    assert T.image_mipmap(str(tmp_path / "a.png"), False, False, 8.0, "repeat", 1.0, True) is a
    assert T.image_mipmap(str(tmp_path / "a.png"), False, True, 8.0, "repeat", 1.0, True) is not a
    assert (a.width, a.height) == (8, 4)


# ---- camera ray differentials ----------------------------------------------------------------------------------------------------
def test_camera_ray_differentials(pkg, oracle):
    H = pkg.host
    film = H.Film(64, 48, "box")
    cam = H.PerspectiveCamera(film, H.Transform.look_at((1, 2, 5), (0, 0.3, 0), (0, 1, 0)).inverse(), fov=40.0)
    base = oracle.generate_ray_differential(cam, (20.25, 30.5))
    # without a lens the auxiliary rays ARE the camera rays of the neighbouring raster positions (perspective.rs:166-171)
    rx = oracle.generate_ray_differential(cam, (21.25, 30.5))
    ry = oracle.generate_ray_differential(cam, (20.25, 31.5))
    assert np.allclose(base[2], base[0]) and np.allclose(base[4], base[0])
    assert np.allclose(base[3], rx[1], atol=1e-6) and np.allclose(base[5], ry[1], atol=1e-6)
    # Ray::scale_differential(1 / sqrt(spp)) (integrator.rs:341): offsets shrink by the factor, the main ray does not move
    s4 = oracle.generate_ray_differential(cam, (20.25, 30.5), spp=4)
    assert np.array_equal(s4[:2], base[:2])
    assert np.allclose(s4[3] - s4[1], 0.5 * (base[3] - base[1]), atol=1e-7) and np.allclose(s4[5] - s4[1], 0.5 * (base[5] - base[1]), atol=1e-7)
    # with a lens: all three rays leave the same lens point and meet their own point of the plane of focus
    lens = H.PerspectiveCamera(film, H.Transform.look_at((1, 2, 5), (0, 0.3, 0), (0, 1, 0)).inverse(), fov=40.0, lensradius=0.2, focaldistance=4.0)
    r = oracle.generate_ray_differential(lens, (20.25, 30.5), plens=(0.7, 0.2)).astype(np.float64)
    pin = oracle.generate_ray_differential(lens, (20.25, 30.5), plens=(0.5, 0.5)).astype(np.float64)  # centre of the lens
    assert np.allclose(r[2], r[0], atol=1e-5) and np.allclose(r[4], r[0], atol=1e-5) and not np.allclose(r[0], pin[0], atol=1e-3)
    fwd = np.array([0, 0.3, 0]) - np.array([1, 2, 5]); fwd /= np.linalg.norm(fwd)

    def focus(o, d):  # where the ray meets the plane at camera-space z = focal distance
        t = (4.0 - (o - np.array([1.0, 2.0, 5.0])) @ fwd) / (d @ fwd)
        return o + t * d

    assert np.allclose(focus(r[0], r[1]), focus(pin[0], pin[1]), atol=1e-4)
    assert np.allclose(focus(r[2], r[3]), focus(pin[2], pin[3]), atol=1e-4)


# ---- the C ABI rejects malformed texture tables (no device needed: the checks run before anything is staged) -----------------------
def test_scene_create_validates_texture_tables(pkg, T):
    import ctypes as C
    H = pkg.host
    lib = pkg.load_library()

    def create(flat):
        d = flat.desc()
        out = C.c_void_p()
        rc = lib.pbrt_b200_scene_create(C.byref(d), 0, C.byref(out))
        msg = lib.pbrt_b200_last_error().decode()
        if rc == 0:
            lib.pbrt_b200_scene_destroy(out)
        return rc, msg

    def scene(tex):
        b = H.SceneBuilder()
        b.material("matte", Kd=tex)
        b.shape("sphere", radius=1.0)
        return b.world_end()

    mip = T.MipMap(np.full((4, 4, 3), 0.5, f32))
    good = scene(T.Tex.scale(T.Tex.imagemap(T.Mapping2D.uv(), mip), T.Tex.checkerboard(T.Mapping2D.uv(), 0.2, 0.8)))
    rc, msg = create(good)
    assert rc in (0, 2), msg  # fine (GPU box) or "no CUDA device" (here) -- never "invalid"
    bad = scene(T.Tex.scale(T.Tex.imagemap(T.Mapping2D.uv(), mip), 0.5))
    bad.textures[-1]["kind"] = T.TEX_MIX  # pops three operands, two are there
    rc, msg = create(bad)
    assert rc == 1 and "texture program" in msg
    bad = scene(T.Tex.imagemap(T.Mapping2D.uv(), mip))
    bad.textures[0]["image"] = 3
    rc, msg = create(bad)
    assert rc == 1 and "image index" in msg
    bad = scene(T.Tex.imagemap(T.Mapping2D.uv(), mip))
    bad.mipmaps[0]["n_levels"] = 9
    rc, msg = create(bad)
    assert rc == 1 and "n_levels" in msg
    bad = scene(T.Tex.imagemap(T.Mapping2D.uv(), mip))
    bad.material_ext[0]["s_tex"][0]["count"] = 7  # runs past the table
    rc, msg = create(bad)
    assert rc == 1 and "texture program" in msg
    bad = scene(T.Tex.constant(0.5))
    bad.material_ext = None  # a `textured` row without its parameters
    rc, msg = create(bad)
    assert rc == 1 and "material_ext" in msg
