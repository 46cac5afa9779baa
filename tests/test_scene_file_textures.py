"""The `Texture` directive and texture-valued material parameters of the scene-file front end (api.rs:1329-1389, make_*_texture :656-704, the
create_* functions of src/textures/*.rs) against the same trees built through the Python API: every texture class, every 2D mapping, parameter
names and defaults, float vs spectrum variants, named materials, shape-level overrides, constant folding, the quirks kept on purpose."""
import importlib
import warnings

import numpy as np
import pytest
from PIL import Image

pkg = importlib.import_module("pbrt-rust_b200")
T = importlib.import_module("pbrt-rust_b200.textures")
H = pkg.host
f32 = np.float32


def _programs(flat, mat_index):
    """The parameter programs of material row `mat_index` as lists of node rows (mipmap indices resolved to pyramid shapes)."""
    x = flat.material_ext[mat_index]

    def prog(ref):
        rows = flat.textures[int(ref["first"]): int(ref["first"]) + int(ref["count"])]
        out = []
        for r in rows:
            img = None
            if int(r["kind"]) == T.TEX_IMAGEMAP:
                m = flat.mipmaps[int(r["image"])]
                img = (int(m["width"]), int(m["height"]), int(m["channels"]), int(m["wrap"]), int(m["do_trilinear"]), float(m["max_anisotropy"]))
            out.append((int(r["kind"]), int(r["mapping"]), int(r["flags"]), np.round(r["v"].astype(np.float64), 6).tolist(), np.round(r["m"].astype(np.float64), 5).tolist(), img))
        return out

    return {"s": [prog(x["s_tex"][k]) for k in range(5)], "f": [prog(x["f_tex"][k]) for k in range(3)], "bump": prog(x["bump"]),
            "s_const": np.round(x["s_const"].astype(np.float64), 6).tolist(), "f_const": np.round(x["f_const"].astype(np.float64), 6).tolist()}


def _built(material, **kw):
    b = H.SceneBuilder()
    b.material(material, **kw)
    b.shape("sphere", radius=1.0)
    return b.world_end()


def _parsed(text, tmp_path):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        api = pkg.pbrt_parse_string("WorldBegin\n" + text + "\nShape \"sphere\"\nWorldEnd", search_dir=str(tmp_path))
    assert not api.errors, api.errors
    return api.jobs[0].flat


def test_every_texture_class_parses_to_the_tree_the_api_builds(tmp_path):
    img = (np.arange(6 * 10 * 3).reshape(6, 10, 3) * 7 % 256).astype(np.uint8)
    Image.fromarray(img).save(tmp_path / "tex.png")
    ident = H.Transform()
    tr = H.Transform.translate((1.0, 2.0, 3.0)) * H.Transform.scale(2.0, 2.0, 2.0)
    M2 = T.Mapping2D
    text = '''
    Texture "a" "color" "imagemap" "string filename" "tex.png" "float uscale" 4 "float vscale" 3 "float udelta" .5 "string wrap" "clamp" "bool trilinear" "true" "float scale" 2
    Texture "b" "float" "imagemap" "string filename" "tex.png" "float maxanisotropy" 4 "string wrap" "black" "bool gamma" "false"
    Texture "c" "color" "checkerboard" "texture tex1" "a" "rgb tex2" [.1 .2 .3] "string aamode" "closedform" "string mapping" "planar" "vector v1" [0 0 2] "vector v2" [0 3 0] "float udelta" .25
    TransformBegin
      Translate 1 2 3
      Scale 2 2 2
      Texture "d" "color" "checkerboard" "integer mapping" 3 "texture tex1" "c" "rgb tex2" [.9 .8 .7]
      Texture "m" "color" "marble" "integer octaves" 5 "float roughness" .4 "float scale" 1.5 "float variation" .3
      Texture "w" "float" "wrinkled" "integer octaves" 3
      Texture "f" "float" "fbm" "float roughness" .7
      Texture "y" "float" "windy"
      Texture "s" "color" "uv" "string mapping" "spherical"
      Texture "q" "float" "dots" "string mapping" "cylindrical" "float inside" .2 "float outside" .8
    TransformEnd
    Texture "x" "color" "mix" "texture tex1" "d" "texture tex2" "m" "texture amount" "f"
    Texture "k" "float" "scale" "texture tex1" "w" "float tex2" .05
    Texture "bl" "color" "bilerp" "rgb v00" [1 0 0] "rgb v11" [0 0 1]
    Texture "cf" "float" "scale" "float tex1" .5 "float tex2" .25
    MakeNamedMaterial "nm" "string type" "uber" "texture Kd" "x" "texture Ks" "s" "texture opacity" "bl" "texture roughness" "q" "texture bumpmap" "k" "float index" 1.3
    NamedMaterial "nm"
    '''
    got = _parsed(text, tmp_path)
    mip_a = T.image_mipmap(str(tmp_path / "tex.png"), False, True, 8.0, "clamp", 2.0, True)
    a = T.Tex.imagemap(M2.uv(4.0, 3.0, 0.5, 0.0), mip_a)
    c = T.Tex.checkerboard(M2.planar((0, 0, 2), (0, 3, 0), 0.25, 0.0), a, np.array([0.1, 0.2, 0.3], f32), "closedform")
    d = T.Tex.checkerboard3d(tr, c, np.array([0.9, 0.8, 0.7], f32))
    m = T.Tex.marble(tr, 5, 0.4, 1.5, 0.3)
    x = T.Tex.mix(d, m, T.Tex.fbm(tr, 8, 0.7))
    s = T.Tex.uv(M2.spherical(tr.inverse()))
    q = T.Tex.dots(M2.cylindrical(tr.inverse()), outside=0.2, inside=0.8)  # create_dots_float hands (inside, outside) to new(map, outside, inside)
    k = T.Tex.scale(T.Tex.wrinkled(tr, 3, 0.5), 0.05)
    bl = T.Tex.bilerp(M2.uv(), (1, 0, 0), 1.0, 0.0, (0, 0, 1))
    want = _built("uber", Kd=x, Ks=s, opacity=bl, roughness=q, bumpmap=k, index=1.3)
    assert got.materials["type"].tolist() == [H.MAT_UBER] and got.materials["textured"].tolist() == [1]
    assert _programs(got, 0) == _programs(want, 0)
    assert (int(got.mipmaps[0]["width"]), int(got.mipmaps[0]["height"])) == (16, 8)  # 10 x 6 resampled to powers of two
    assert np.allclose(got.mipmap_objects[0].texels, want.mipmap_objects[0].texels)


def test_float_image_maps_constant_folding_and_shape_level_overrides(tmp_path):
    Image.fromarray(np.full((4, 4, 3), 128, np.uint8)).save(tmp_path / "g.png")
    text = '''
    Texture "b" "float" "imagemap" "string filename" "g.png" "float maxanisotropy" 4 "string wrap" "black" "bool gamma" "false"
    Texture "cf" "float" "scale" "float tex1" .5 "float tex2" .25
    Texture "cm" "color" "mix" "rgb tex1" [1 0 0] "rgb tex2" [0 0 1] "float amount" .25
    Texture "cc" "color" "constant" "rgb value" [.3 .4 .5]
    Material "plastic" "texture Kd" "cm" "texture roughness" "cf"
    '''
    got = _parsed(text, tmp_path)
    # every parameter folded to a constant: a plain row on the fast kernels
    assert got.materials["textured"].tolist() == [0] and got.material_ext is None
    assert np.allclose(got.materials[0]["a"], [0.75, 0.0, 0.25]) and got.materials[0]["f0"] == f32(0.125)
    # a shape-level parameter overrides the material's (get_materialfor_shape, api.rs:363-380): the sphere takes the image map as sigma
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        api = pkg.pbrt_parse_string('WorldBegin\nTexture "b" "float" "imagemap" "string filename" "g.png" "string wrap" "black" "bool gamma" "false"\n'
                                    'Texture "cc" "color" "constant" "rgb value" [.3 .4 .5]\nMaterial "matte" "texture Kd" "cc"\nShape "sphere"\n'
                                    'Shape "sphere" "texture sigma" "b" "float radius" 2\nWorldEnd', search_dir=str(tmp_path))
    flat = api.jobs[0].flat
    assert flat.materials["textured"].tolist() == [0, 1] and np.allclose(flat.materials[0]["a"], [0.3, 0.4, 0.5])
    p = _programs(flat, 1)
    assert p["s_const"][0] == [0.3, 0.4, 0.5] and len(p["f"][0]) == 1 and p["f"][0][0][0] == T.TEX_IMAGEMAP
    assert p["f"][0][0][5] == (4, 4, 1, T.WRAP["black"], 0, 8.0)
    assert np.allclose(flat.mipmap_objects[0].texels[:16], 128 / 255.0, rtol=1e-6)  # luminance of a grey texel, no gamma


def test_texture_directive_quirks_and_errors(tmp_path):
    # float variants the reference does not have (create_uv_float / create_marble_float return None): the texture is not defined,
    # and a material that names it falls back to its default with a warning
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        flat = pkg.pbrt_parse_string('WorldBegin\nTexture "u" "float" "uv"\nTexture "m" "float" "marble"\nMaterial "matte" "texture sigma" "u"\nShape "sphere"\nWorldEnd').jobs[0].flat
    assert flat.materials["textured"].tolist() == [0] and flat.materials[0]["f0"] == 0.0
    assert any("Couldn't find float texture" in str(x.message) for x in w)
    # unknown class / unknown type
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        api = pkg.pbrt_parse_string('WorldBegin\nTexture "z" "color" "nosuch"\nTexture "z2" "vector" "constant"\nWorldEnd')
    assert any('Spectrum texture "nosuch" unknown' in str(x.message) for x in w) and any("unknown" in e for e in api.errors)
    # a checkerboard dimension other than 2 or 3
    api = pkg.pbrt_parse_string('WorldBegin\nTexture "z" "color" "checkerboard" "integer mapping" 4\nWorldEnd')
    assert any("dimensional checkerboard" in e for e in api.errors)
    # an unknown antialiasing mode means closedform (checkerboard.rs:122-131)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        flat = _parsed('Texture "c" "color" "checkerboard" "string aamode" "fancy"\nMaterial "matte" "texture Kd" "c"', tmp_path)
    assert int(flat.textures[-1]["flags"]) == T.TEX_AA_CLOSEDFORM
    # substrate / metal / glass / mirror slots
    flat = _parsed('Texture "c" "color" "checkerboard"\nTexture "f" "float" "checkerboard" "float tex1" .1 "float tex2" .3\n'
                   'Material "substrate" "texture Kd" "c" "texture vroughness" "f" "bool remaproughness" "false"', tmp_path)
    p = _programs(flat, 0)
    assert flat.materials[0]["type"] == H.MAT_SUBSTRATE and flat.materials[0]["remap_roughness"] == 0
    assert len(p["s"][0]) == 3 and p["s_const"][1] == [0.5, 0.5, 0.5] and p["f_const"][0] == 0.1 and len(p["f"][1]) == 3
    flat = _parsed('Texture "f" "float" "checkerboard" "float tex1" .1 "float tex2" .3\nMaterial "metal" "texture roughness" "f"', tmp_path)
    p = _programs(flat, 0)
    assert len(p["f"][0]) == 3 and p["f"][0] == p["f"][1]  # metal.rs:88-97: u / v roughness fall back to `roughness`
