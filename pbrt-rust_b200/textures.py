"""Host side of the texture system: texture expression trees, MIPMap pyramids, image files.

The reference keeps textures as `Arc<Textures>` trees (src/core/texture.rs:98-120) that materials evaluate per hit.  Here a tree is a
`Tex` object; `flatten()` turns it into the postfix `pbrt_b200_texnode` program the device (and the CPU oracle) walk, and `MipMap`
builds the pyramid of an image map the way `MIPMap::new` does (src/core/mipmap.rs:76-198: Lanczos resampling to a power of two, then
2x2 box-filtered levels).  Nothing here evaluates a texture at a hit; that happens in csrc/texture.cuh.
"""
from __future__ import annotations

import warnings

import numpy as np

f32 = np.float32

TEXNODE_DTYPE = np.dtype([("kind", "<u4"), ("mapping", "<u4"), ("flags", "<u4"), ("image", "<u4"), ("v", "<f4", 12), ("m", "<f4", 16)])  # pbrt_b200_texnode
MIPMAP_DTYPE = np.dtype([("texels", "<u8"), ("n_levels", "<u4"), ("channels", "<u4"), ("width", "<u4"), ("height", "<u4"), ("wrap", "<u4"),
                         ("do_trilinear", "<u4"), ("max_anisotropy", "<f4"), ("pad", "<u4")])  # pbrt_b200_mipmap
TEXREF_DTYPE = np.dtype([("first", "<u4"), ("count", "<u4")])
MATERIAL_EXT_DTYPE = np.dtype([("s_tex", TEXREF_DTYPE, 5), ("s_const", "<f4", (5, 3)), ("f_tex", TEXREF_DTYPE, 3), ("f_const", "<f4", 3),
                               ("bump", TEXREF_DTYPE), ("pad", "<u4", 4)])  # pbrt_b200_material_ext
assert TEXNODE_DTYPE.itemsize == 128 and MIPMAP_DTYPE.itemsize == 40 and MATERIAL_EXT_DTYPE.itemsize == 160

(TEX_CONSTANT, TEX_SCALE, TEX_MIX, TEX_BILERP, TEX_IMAGEMAP, TEX_UV, TEX_CHECKERBOARD2D, TEX_CHECKERBOARD3D, TEX_DOTS, TEX_FBM, TEX_WRINKLED,
 TEX_MARBLE, TEX_WINDY) = range(13)
MAP_UV, MAP_SPHERICAL, MAP_CYLINDRICAL, MAP_PLANAR = range(4)
TEX_AA_CLOSEDFORM = 1
WRAP = {"repeat": 0, "black": 1, "clamp": 2}


# ------------------------------------------------------------------------------------------------------------------
# 2D mappings, texture.rs:122-283 (get_mapping2d :433-463)
# ------------------------------------------------------------------------------------------------------------------
class Mapping2D:
    def __init__(self, kind, m):
        self.kind = kind
        self.m = np.zeros(16, f32)
        m = np.asarray(m, f32).reshape(-1)
        self.m[: len(m)] = m

    @staticmethod
    def uv(su=1.0, sv=1.0, du=0.0, dv=0.0):
        return Mapping2D(MAP_UV, [su, sv, du, dv])

    @staticmethod
    def planar(vs=(1, 0, 0), vt=(0, 1, 0), ds=0.0, dt=0.0):
        return Mapping2D(MAP_PLANAR, [*vs, *vt, ds, dt])

    @staticmethod
    def spherical(world_to_texture):  # a host.Transform
        return Mapping2D(MAP_SPHERICAL, world_to_texture.m)

    @staticmethod
    def cylindrical(world_to_texture):
        return Mapping2D(MAP_CYLINDRICAL, world_to_texture.m)

    def key(self):
        return (self.kind, self.m.tobytes())


# ------------------------------------------------------------------------------------------------------------------
# texture trees
# ------------------------------------------------------------------------------------------------------------------
def _as_rgb(v):
    return np.full(3, v, f32) if np.isscalar(v) or np.ndim(v) == 0 else np.asarray(v, f32).reshape(3)


class Tex:
    """One node of a texture tree.  `children` are Tex objects or constants (f32 scalar / RGB triple)."""

    def __init__(self, kind, children=(), mapping=None, flags=0, v=(), m=None, image=None):
        self.kind, self.children, self.mapping, self.flags, self.image = kind, tuple(children), mapping, flags, image
        self.v = np.zeros(12, f32)
        v = np.asarray(v, f32).reshape(-1)
        self.v[: len(v)] = v
        self.m = np.zeros(16, f32) if m is None else np.asarray(m, f32).reshape(16)

    def key(self):
        kids = tuple(c.key() if isinstance(c, Tex) else _as_rgb(c).tobytes() for c in self.children)
        return (self.kind, kids, self.mapping.key() if self.mapping else None, self.flags, self.v.tobytes(), self.m.tobytes(), id(self.image))

    # constructors named after the reference's create_* functions ------------------------------------------------
    @staticmethod
    def constant(value):  # textures/constant.rs
        return Tex(TEX_CONSTANT, v=_as_rgb(value))

    @staticmethod
    def scale(tex1, tex2):  # textures/scaled.rs
        return Tex(TEX_SCALE, (tex1, tex2))

    @staticmethod
    def mix(tex1, tex2, amount):  # textures/mix.rs
        return Tex(TEX_MIX, (tex1, tex2, amount))

    @staticmethod
    def bilerp(mapping, v00, v01, v10, v11):  # textures/biler.rs
        return Tex(TEX_BILERP, mapping=mapping, v=np.concatenate([_as_rgb(x) for x in (v00, v01, v10, v11)]))

    @staticmethod
    def imagemap(mapping, mipmap):  # textures/imagemap.rs
        return Tex(TEX_IMAGEMAP, mapping=mapping, image=mipmap)

    @staticmethod
    def uv(mapping):  # textures/uv.rs
        return Tex(TEX_UV, mapping=mapping)

    @staticmethod
    def checkerboard(mapping, tex1=1.0, tex2=0.0, aamode="none"):  # textures/checkerboard.rs:28-73
        return Tex(TEX_CHECKERBOARD2D, (tex1, tex2), mapping=mapping, flags=0 if aamode == "none" else TEX_AA_CLOSEDFORM)

    @staticmethod
    def checkerboard3d(world_to_texture, tex1=1.0, tex2=0.0):  # textures/checkerboard.rs:88-100
        return Tex(TEX_CHECKERBOARD3D, (tex1, tex2), m=world_to_texture.m)

    @staticmethod
    def dots(mapping, outside, inside):  # textures/dots.rs (struct field order; create_dots_* passes (inside, outside) into it)
        return Tex(TEX_DOTS, (outside, inside), mapping=mapping)

    @staticmethod
    def fbm(world_to_texture, octaves=8, omega=0.5):  # textures/fbm.rs
        return Tex(TEX_FBM, v=[omega, octaves], m=world_to_texture.m)

    @staticmethod
    def wrinkled(world_to_texture, octaves=8, omega=0.5):  # textures/wrinkled.rs
        return Tex(TEX_WRINKLED, v=[omega, octaves], m=world_to_texture.m)

    @staticmethod
    def marble(world_to_texture, octaves=8, omega=0.5, scale=1.0, variation=0.2):  # textures/marble.rs
        return Tex(TEX_MARBLE, v=[omega, octaves, scale, variation], m=world_to_texture.m)

    @staticmethod
    def windy(world_to_texture):  # textures/windy.rs
        return Tex(TEX_WINDY, m=world_to_texture.m)


def is_texture(v):
    return isinstance(v, Tex)


class TextureTables:
    """Accumulates the flat `textures[]` / `mipmaps[]` arrays of a scene."""

    def __init__(self):
        self.nodes = []     # TEXNODE_DTYPE rows
        self.mipmaps = []   # MipMap objects, in table order
        self._mip_index = {}
        self._programs = {}  # Tex.key() -> (first, count)
        self.max_depth = 0   # deepest value stack any program needs

    def _emit(self, t):
        if not isinstance(t, Tex):
            t = Tex.constant(t)
        for c in t.children:
            self._emit(c)
        r = np.zeros(1, TEXNODE_DTYPE)[0]
        r["kind"], r["flags"], r["v"] = t.kind, t.flags, t.v
        if t.mapping is not None:
            r["mapping"], r["m"] = t.mapping.kind, t.mapping.m
        else:
            r["m"] = t.m
        if t.image is not None:
            if id(t.image) not in self._mip_index:
                self._mip_index[id(t.image)] = len(self.mipmaps)
                self.mipmaps.append(t.image)
            r["image"] = self._mip_index[id(t.image)]
        self.nodes.append(r)

    def program(self, tex):
        """-> (first, count) of the postfix program of `tex` (shared between equal trees)."""
        k = tex.key()
        if k not in self._programs:
            first = len(self.nodes)
            self._emit(tex)
            self._programs[k] = (first, len(self.nodes) - first)
            self.max_depth = max(self.max_depth, max_stack_depth(self.nodes[first:]))
        return self._programs[k]

    def node_array(self):
        return np.array(self.nodes, TEXNODE_DTYPE) if self.nodes else np.zeros(0, TEXNODE_DTYPE)

    def mipmap_array(self):
        out = np.zeros(len(self.mipmaps), MIPMAP_DTYPE)
        for i, m in enumerate(self.mipmaps):
            out[i]["texels"] = m.texels.ctypes.data
            out[i]["n_levels"], out[i]["channels"], out[i]["width"], out[i]["height"] = m.n_levels, m.channels, m.width, m.height
            out[i]["wrap"], out[i]["do_trilinear"], out[i]["max_anisotropy"] = m.wrap, 1 if m.do_trilinear else 0, m.max_anisotropy
        return out


def max_stack_depth(nodes):
    """Value-stack depth the postfix walk over ONE program needs (the builder checks it against the device's stack)."""
    pops = {TEX_SCALE: 2, TEX_MIX: 3, TEX_CHECKERBOARD2D: 2, TEX_CHECKERBOARD3D: 2, TEX_DOTS: 2}
    depth = best = 0
    for n in nodes:
        depth += 1 - pops.get(int(n["kind"]), 0)
        best = max(best, depth)
    return best


# ------------------------------------------------------------------------------------------------------------------
# MIPMap::new, mipmap.rs:76-198
# ------------------------------------------------------------------------------------------------------------------
def _lanczos(x, tau):  # texture.rs:319-328
    x = np.abs(x).astype(f32)
    xp = (x * f32(np.pi)).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (np.sin(xp * f32(tau)).astype(f32) / (xp * f32(tau)).astype(f32)).astype(f32)
        lanc = (np.sin(xp).astype(f32) / xp).astype(f32)
    out = (s * lanc).astype(f32)
    out = np.where(x > f32(1.0), f32(0.0), out)
    return np.where(x < f32(1.0e-5), f32(1.0), out).astype(f32)


def resample_weights(oldres, newres):  # mipmap.rs:275-299
    assert newres >= oldres
    i = np.arange(newres, dtype=f32)
    center = ((i + f32(0.5)) * f32(oldres) / f32(newres)).astype(f32)
    first = np.floor((center - f32(2.0)) + f32(0.5)).astype(np.int64)
    w = np.zeros((newres, 4), f32)
    for j in range(4):
        pos = (first.astype(f32) + f32(j) + f32(0.5)).astype(f32)
        w[:, j] = _lanczos(((pos - center) / f32(2.0)).astype(f32), 2.0)
    inv = (f32(1.0) / (w[:, 0] + w[:, 1] + w[:, 2] + w[:, 3]).astype(f32)).astype(f32)
    return first, (w * inv[:, None]).astype(f32)


def _wrap_index(i, n, wrap):
    """-> (index, valid) for texel coordinate array i under `wrap` (mipmap.rs:301-321; Clamp clamps to n - 1: the reference's
    clamp(s, 0, u) would index one past the row, a panic nobody can mirror)."""
    if wrap == WRAP["repeat"]:
        return np.mod(i, n), np.ones(i.shape, bool)
    if wrap == WRAP["clamp"]:
        return np.clip(i, 0, n - 1), np.ones(i.shape, bool)
    ok = (i >= 0) & (i < n)
    return np.where(ok, i, 0), ok


def _round_up_pow2(v):
    return 1 << max(int(v) - 1, 0).bit_length()


class MipMap:
    """The image pyramid of an ImageTexture (texels: (h, w, c) float32, already flipped / converted, imagemap.rs:144-160)."""

    def __init__(self, texels, do_trilinear=False, max_anisotropy=8.0, wrap="repeat"):
        img = np.ascontiguousarray(texels, f32)
        if img.ndim == 2:
            img = img[:, :, None]
        h, w, c = img.shape
        assert c in (1, 3)
        self.wrap = WRAP[wrap] if isinstance(wrap, str) else int(wrap)
        self.do_trilinear, self.max_anisotropy, self.channels = bool(do_trilinear), float(max_anisotropy), c
        if (w & (w - 1)) or (h & (h - 1)):
            img = self._resample(img, _round_up_pow2(w), _round_up_pow2(h))
            h, w = img.shape[:2]
        self.width, self.height = w, h
        levels = [img]
        n_levels = 1 + int(np.log2(f32(max(w, h))))
        for _ in range(1, n_levels):
            levels.append(self._halve(levels[-1]))
        self.levels = levels
        self.n_levels = n_levels
        self.texels = np.ascontiguousarray(np.concatenate([l.reshape(-1) for l in levels]).astype(f32))

    def _resample(self, img, nw, nh):
        h, w, c = img.shape
        first, wt = resample_weights(w, nw)
        tmp = np.zeros((h, nw, c), f32)
        for j in range(4):  # row[s] += data[t * res.x + origs] * weight[j], j in order
            idx, ok = _wrap_index(first + j, w, self.wrap)
            tmp = (tmp + np.where(ok[None, :, None], img[:, idx, :] * wt[None, :, j, None], f32(0.0)).astype(f32)).astype(f32)
        first, wt = resample_weights(h, nh)
        out = np.zeros((nh, nw, c), f32)
        for j in range(4):
            idx, ok = _wrap_index(first + j, h, self.wrap)
            out = (out + np.where(ok[:, None, None], tmp[idx, :, :] * wt[:, j, None, None], f32(0.0)).astype(f32)).astype(f32)
        return np.clip(out, f32(0.0), f32(np.inf)).astype(f32)  # Clampable::clamp(work_data[t], 0, INFINITY)

    def _halve(self, prev):
        ph, pw, _ = prev.shape
        sres, tres = max(1, pw // 2), max(1, ph // 2)
        s, t = np.arange(sres), np.arange(tres)

        def texel(si, ti):
            ix, okx = _wrap_index(si, pw, self.wrap)
            iy, oky = _wrap_index(ti, ph, self.wrap)
            v = prev[iy[:, None], ix[None, :], :]
            return np.where((oky[:, None] & okx[None, :])[:, :, None], v, f32(0.0)).astype(f32)

        a = texel(2 * s, 2 * t)
        b = texel(2 * s + 1, 2 * t)
        cc = texel(2 * s, 2 * t + 1)
        d = texel(2 * s + 1, 2 * t + 1)
        return ((((a + b).astype(f32) + cc).astype(f32) + d).astype(f32) * f32(0.25)).astype(f32)


# ------------------------------------------------------------------------------------------------------------------
# image files: read_image, src/core/imageio.rs:22-38
# ------------------------------------------------------------------------------------------------------------------
def read_pfm(path):  # imageio.rs:168-260: "PF" / "Pf", width height, scale (sign = endianness), rows bottom-up in the file
    with open(path, "rb") as f:
        data = f.read()
    pos = 0

    def word():
        nonlocal pos
        s = bytearray()
        while pos < len(data):
            c = data[pos]
            pos += 1
            if c in b" \n\t":
                break
            s.append(c)
        return s.decode("ascii", "replace")

    magic = word()
    if magic not in ("PF", "Pf"):
        raise ValueError(f'Error reading PFM file "{path}"')
    nc = 3 if magic == "PF" else 1
    w, h = int(word()), int(word())
    scale = float(word())
    raw = np.frombuffer(data, dtype="<f4" if scale < 0 else ">f4", count=w * h * nc, offset=pos).astype(f32)
    raw = (raw * f32(abs(scale))).astype(f32).reshape(h, w, nc)[::-1]  # flip in y: the reference's reader stores the top row first
    if nc == 1:
        raw = np.repeat(raw, 3, axis=2)
    return np.ascontiguousarray(raw)


def read_image(path):
    """-> (h, w, 3) float32 RGB, top row first (what `read_image` returns before ImageTexture flips it)."""
    low = str(path).lower()
    if low.endswith(".pfm"):
        return read_pfm(path)
    if low.endswith((".png", ".tga")):  # read_image_png_tga, imageio.rs:338-357: to_rgb8, then u8 / 255
        from PIL import Image

        im = np.asarray(Image.open(path).convert("RGB"), np.uint8)
        return (im.astype(f32) / f32(255.0)).astype(f32)
    if low.endswith((".exr", ".hdr")):
        import os

        os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
        import cv2

        im = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
        if im is None:
            raise ValueError(f'cannot read "{path}"')
        if im.ndim == 2:
            im = np.repeat(im[:, :, None], 3, axis=2)
        return np.ascontiguousarray(im[:, :, 2::-1].astype(f32))
    raise ValueError(f'Unable to load image stored in format "{low.rsplit(".", 1)[-1]}" for filename "{path}"')


def inverse_gamma_correct(v):  # pbrt.rs:218-222
    v = np.asarray(v, f32)
    with np.errstate(invalid="ignore"):
        hi = np.power(((v + f32(0.055)).astype(f32) * f32(1.0) / f32(1.055)).astype(f32), f32(2.4)).astype(f32)
    return np.where(v <= f32(0.04045), (v * f32(1.0) / f32(12.92)).astype(f32), hi).astype(f32)


def convert_texels(rgb, as_float, scale, gamma):
    """ImageTexture::get_texture, imagemap.rs:144-160: flip in y, then convert_from (:71-98)."""
    rgb = np.ascontiguousarray(rgb[::-1], f32)
    if as_float:  # Float::convert_from: scale * (gamma ? inverse_gamma_correct(y) : y)
        y = ((f32(0.212671) * rgb[..., 0] + f32(0.715160) * rgb[..., 1]).astype(f32) + f32(0.072169) * rgb[..., 2]).astype(f32)
        g = inverse_gamma_correct(y) if gamma else y
        return (f32(scale) * g).astype(f32)[:, :, None]
    g = inverse_gamma_correct(rgb) if gamma else rgb
    return (g * f32(scale)).astype(f32)


_MIP_CACHE = {}


def image_mipmap(filename, as_float, do_trilinear, max_aniso, wrap, scale, gamma):
    """ImageTexture::get_texture with its cache keyed like TexInfo (imagemap.rs:36-62,109-162)."""
    key = (str(filename), bool(do_trilinear), float(max_aniso), float(scale), bool(gamma), wrap, bool(as_float))
    if key not in _MIP_CACHE:
        try:
            rgb = read_image(filename)
        except Exception as e:  # noqa: BLE001 -- the reference logs and substitutes (imagemap.rs:136-142)
            warnings.warn(f'Creating a constant grey texture to replace "{filename}". ({e})')
            rgb = np.full((1, 1, 3), 0.5, f32)
        _MIP_CACHE[key] = MipMap(convert_texels(rgb, as_float, scale, gamma), do_trilinear, max_aniso, wrap)
    return _MIP_CACHE[key]


def clear_cache():
    _MIP_CACHE.clear()
