"""ParamSet / TextureParams of pbrt-rust (src/core/paramset.rs), host side only.

A `ParamSet` is what every scene-file directive carries (`"float fov" [30]`, `"rgb Kd" [.5 .5 .5]` ...).
Lookups follow the reference exactly: `find_one_*` only matches an item holding ONE value
(paramset.rs:39-51), `find_*` returns the whole array (:24-37), spectra are converted to RGB when
they are added (:131-250), adding a name twice replaces the earlier item (`erase_*`).
"""
from __future__ import annotations

import os
import warnings

import numpy as np

from . import spectrum as S

f32 = np.float32

_search_dir = None
_cached_spectra = {}


def set_search_directory(d):  # fileutil.rs: SEARCH_DIR
    global _search_dir
    _search_dir = d


def resolve_filename(name):  # fileutil.rs:42-60
    if _search_dir is None or not name or os.path.isabs(name):
        return name
    return os.path.join(_search_dir, *name.split("/"))


TYPE_ALIASES = {"int": "int", "integer": "int", "bool": "bool", "float": "float", "vector2": "vector2", "vector3": "vector3", "vector": "vector3",
                "point2": "point2", "point3": "point3", "point": "point3", "normal": "normal", "rgb": "rgb", "color": "rgb", "xyz": "xyz",
                "blackbody": "blackbody", "spectrum": "spectrum", "string": "string", "texture": "texture"}


class ParamSet:
    """Typed buckets keyed by name; values are numpy arrays (numeric) or lists (bool / string)."""

    BUCKETS = ("bools", "ints", "floats", "point2fs", "vector2fs", "point3fs", "vector3fs", "normals", "spectra", "strings", "textures")

    def __init__(self):
        for b in self.BUCKETS:
            setattr(self, b, {})
        self.looked_up = set()

    # --- add_* (pbrtparser.rs:181-443 -> paramset.rs add_*) ------------------------
    def add(self, decl, values):
        """One `"type name" values` item as the grammar's ParamList produces it (pbrtparser.rs:130-163)."""
        parts = decl.split()
        if len(parts) < 2:
            raise ValueError(f'parameter declaration "{decl}" needs a type and a name')
        typ, name = parts[0], parts[1]
        if typ not in TYPE_ALIASES:
            raise ValueError(f"unknown parameter type {typ}")  # the reference panics
        ty = TYPE_ALIASES[typ]
        is_str = isinstance(values, list)
        if ty in ("texture", "string", "bool"):
            if not is_str:
                warnings.warn(f'Expected string parameter value for parameter "{name}" with type "{ty}"')
                return
        elif ty != "spectrum" and is_str:
            warnings.warn(f'Expected numeric parameter value for parameter "{name}" with type "{ty}"')
            return
        getattr(self, "_add_" + ty)(name, values)

    def _trunc(self, name, v, k, what):
        excess = len(v) % k
        if excess:
            warnings.warn(f'Excess values given with {what} parameter "{name}". Ignoring last {excess} of them.')
            v = v[: len(v) - excess]
        return v

    def _add_int(self, name, v):
        self.ints[name] = np.trunc(np.asarray(v, f32)).astype(np.int64)  # `*x as isize`

    def _add_bool(self, name, v):
        out = []
        for x in v:
            if x not in ("true", "false"):
                warnings.warn(f'Value "{x}" unknown for Boolean Parameter "{name}". Using "false"')
            out.append(x == "true")
        self.bools[name] = out

    def _add_float(self, name, v):
        self.floats[name] = np.asarray(v, f32)

    def _add_point2(self, name, v):
        self.point2fs[name] = self._trunc(name, np.asarray(v, f32), 2, "point2").reshape(-1, 2)

    def _add_vector2(self, name, v):
        self.vector2fs[name] = self._trunc(name, np.asarray(v, f32), 2, "vector2").reshape(-1, 2)

    def _add_point3(self, name, v):
        self.point3fs[name] = self._trunc(name, np.asarray(v, f32), 3, "point3").reshape(-1, 3)

    def _add_vector3(self, name, v):
        self.vector3fs[name] = self._trunc(name, np.asarray(v, f32), 3, "vector3").reshape(-1, 3)

    def _add_normal(self, name, v):
        self.normals[name] = self._trunc(name, np.asarray(v, f32), 3, "normal").reshape(-1, 3)

    def _add_rgb(self, name, v):  # add_rgb_spectrum, paramset.rs:131-145
        self.spectra[name] = self._trunc(name, np.asarray(v, f32), 3, "RGB").reshape(-1, 3)

    def _add_xyz(self, name, v):  # :147-161
        v = self._trunc(name, np.asarray(v, f32), 3, "XYZ").reshape(-1, 3)
        self.spectra[name] = np.array([S.xyz_to_rgb(x) for x in v], f32).reshape(-1, 3)

    def _add_blackbody(self, name, v):  # :163-179: (temperature, scale) pairs
        v = self._trunc(name, np.asarray(v, f32), 2, "blackbody").reshape(-1, 2)
        self.spectra[name] = np.array([S.blackbody_rgb(t, s) for t, s in v], f32).reshape(-1, 3)

    def _add_spectrum(self, name, v):
        if isinstance(v, list):  # add_sampled_spectrum_files, :199-250
            out = []
            for n in v:
                fname = os.path.abspath(resolve_filename(n))
                if fname not in _cached_spectra:
                    try:
                        vals = S.read_float_file(fname)
                    except (OSError, ValueError) as e:
                        warnings.warn(f"{e}: Unable to read SPD file {fname}. Using black distribution")
                        out.append(np.zeros(3, f32))
                        continue
                    # floatfile.rs:22-30 pushes every parsed token TWICE (once in the match, once after `?`), so the
                    # (wavelength, value) pairing below sees w0 w0 v0 v0 w1 w1 ...; kept, it decides the colour.
                    vals = np.repeat(vals, 2)
                    if len(vals) % 2:
                        warnings.warn(f'Extra value found in spectrum file "{fname}". Ignoring it.')
                    m = len(vals) // 2
                    _cached_spectra[fname] = S.from_sampled(vals[0:2 * m:2], vals[1:2 * m:2])
                out.append(_cached_spectra[fname])
            self.spectra[name] = np.array(out, f32).reshape(-1, 3)
            return
        v = self._trunc(name, np.asarray(v, f32), 2, "sampled spectrum").reshape(-1, 2)  # add_sampled_spectrum, :181-197
        self.spectra[name] = S.from_sampled(v[:, 0], v[:, 1]).reshape(1, 3)

    def _add_string(self, name, v):
        self.strings[name] = list(v)

    def _add_texture(self, name, v):
        if len(v) != 1:
            warnings.warn(f'Only one parameter allowed for "texture" parameter "{name}"')
        self.textures[name] = [v[0]]

    # --- find_* ------------------------------------------------------------------------
    def _find(self, bucket, name):
        d = getattr(self, bucket)
        if name in d:
            self.looked_up.add((bucket, name))
            return d[name]
        return None

    def _find_one(self, bucket, name, default):
        d = getattr(self, bucket)
        if name in d and len(d[name]) == 1:
            self.looked_up.add((bucket, name))
            return d[name][0]
        return default

    def find_one_float(self, n, d):
        return f32(self._find_one("floats", n, d))

    def find_one_int(self, n, d):
        return int(self._find_one("ints", n, d))

    def find_one_bool(self, n, d):
        return bool(self._find_one("bools", n, d))

    def find_one_string(self, n, d):
        return self._find_one("strings", n, d)

    def find_one_point3f(self, n, d):
        return np.asarray(self._find_one("point3fs", n, d), f32)

    def find_one_spectrum(self, n, d):
        v = self._find_one("spectra", n, None)
        return np.full(3, d, f32) if v is None and np.isscalar(d) else np.asarray(d if v is None else v, f32)

    def find_one_filename(self, n, d):  # paramset.rs:278-284
        fn = self.find_one_string(n, d)
        return d if not fn else os.path.abspath(resolve_filename(fn))

    def find_texture(self, n, d=""):
        return self._find_one("textures", n, d)

    def find_float(self, n):
        return self._find("floats", n)

    def find_int(self, n):
        return self._find("ints", n)

    def find_point3f(self, n):
        return self._find("point3fs", n)

    def find_point2f(self, n):
        return self._find("point2fs", n)

    def find_vector3f(self, n):
        return self._find("vector3fs", n)

    def find_normal3f(self, n):
        return self._find("normals", n)

    def find_spectrum(self, n):
        return self._find("spectra", n)

    def report_unused(self, where=""):  # paramset.rs:286-310
        for b in self.BUCKETS:
            for name in getattr(self, b):
                if (b, name) not in self.looked_up:
                    warnings.warn(f'Parameter "{name}" not used{where}')


class TextureParams:
    """paramset.rs:443-610: shape parameters shadow the material's; textures resolve by name.

    A texture resolves to a constant (f32 for float textures, RGB for spectrum ones) when its whole expression is constant,
    otherwise to a textures.Tex tree.
    """

    def __init__(self, geo, mat, float_textures, spectrum_textures):
        self.geo, self.mat, self.ftex, self.stex = geo, mat, float_textures, spectrum_textures

    def find_float(self, n, d):
        return self.geo.find_one_float(n, self.mat.find_one_float(n, d))

    def find_string(self, n, d):
        return self.geo.find_one_string(n, self.mat.find_one_string(n, d))

    def find_bool(self, n, d):
        return self.geo.find_one_bool(n, self.mat.find_one_bool(n, d))

    def find_spectrum(self, n, d):
        return self.geo.find_one_spectrum(n, self.mat.find_one_spectrum(n, d))

    def find_int(self, n, d):
        return self.geo.find_one_int(n, self.mat.find_one_int(n, d))

    def find_vector3f(self, n, d):
        return np.asarray(self.geo._find_one("vector3fs", n, self.mat._find_one("vector3fs", n, d)), f32)

    def find_filename(self, n, d):
        return self.geo.find_one_filename(n, self.mat.find_one_filename(n, d))

    def _tex_or_null(self, n, find, table, kind):
        name = self.geo.find_texture(n, "")
        if not name:
            s = find(self.geo, n)
            if s is not None:
                if len(s) > 1:
                    warnings.warn(f'Ignoring excess values provided with parameter "{n}"')
                return s[0]
            name = self.mat.find_texture(n, "")
            if not name:
                s = find(self.mat, n)
                if s is not None:
                    if len(s) > 1:
                        warnings.warn(f'Ignoring excess values provided with parameter "{n}"')
                    return s[0]
            if not name:
                return None
        if name in table:
            return table[name]
        warnings.warn(f'Couldn\'t find {kind} texture named "{name}" for parameter "{n}"')
        return None

    def get_spectrumtexture_ornull(self, n):
        return self._tex_or_null(n, ParamSet.find_spectrum, self.stex, "spectrum")

    def get_floattexture_ornull(self, n):
        return self._tex_or_null(n, ParamSet.find_float, self.ftex, "float")

    def get_spectrumtexture(self, n, d):
        v = self.get_spectrumtexture_ornull(n)
        if v is not None and not isinstance(v, np.ndarray) and not np.isscalar(v):
            return v  # a textures.Tex tree
        return np.asarray(d if v is None else v, f32) if not np.isscalar(d) or v is not None else np.full(3, d, f32)

    def get_floattexture(self, n, d):
        v = self.get_floattexture_ornull(n)
        if v is not None and not np.isscalar(v) and not isinstance(v, np.ndarray):
            return v  # a textures.Tex tree
        return f32(d if v is None else v)

    def report_unused(self):
        self.geo.report_unused()
        self.mat.report_unused()
