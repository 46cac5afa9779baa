"""Scene-file front end: the `.pbrt` lexer and grammar of pbrt-rust (src/pbrtparser/lexer.rs,
src/commands.lalrpop, src/pbrtparser/pbrtparser.rs), driving `api.API` exactly as `pbrtparser::parse`
drives `core::api::API` (pbrtparser.rs:38-91).

    import importlib; pkg = importlib.import_module("pbrt-rust_b200")
    job = pkg.pbrt_parse("scene.pbrt")          # lex, parse, run every directive up to WorldEnd
    image, stats = job.render(device=0)         # the CUDA path; there is no CPU fallback

Token rules follow lexer.rs: numbers are `[+-]?(\\d+(\\.\\d*)?([eE][+-]?\\d+)?|\\.\\d+([eE][+-]?\\d+)?)` parsed
straight to f32 (correctly rounded, as Rust's `str::parse::<f32>`), strings are `"[^"]*"`, comments run from
`#` to end of line, directive keywords are matched as prefixes in the order of lexer.rs:128-173.  Two deliberate
differences, both supersets: text the lexer cannot tokenise is an error here (lexer.rs:121-175 silently ends the
token stream, dropping the rest of the file), and `Identity` / `TransformTimes` — which the reference lexes but
its grammar never accepts (commands.lalrpop:66-102) — are parsed and forwarded to `API.identity` /
`API.transform_times`, which exist in api.rs:992,1122.
"""
from __future__ import annotations

import os
import re
from fractions import Fraction

import numpy as np

from . import paramset as PS
from .host import B200Error

f32 = np.float32

# keyword order of lexer.rs:130-173 (prefix match, first hit wins); TransformTimes moved before Transform so that it lexes at all
KEYWORDS = ["AttributeBegin", "AttributeEnd", "ActiveTransform", "All", "StartTime", "EndTime", "AreaLightSource", "Accelerator", "ConcatTransform",
            "CoordinateSystem", "CoordSysTransform", "Camera", "Film", "Integrator", "Include", "Identity", "LightSource", "LookAt", "Material",
            "MakeNamedMaterial", "MakeNamedMedium", "NamedMaterial", "MediumInterface", "ObjectBegin", "ObjectEnd", "ObjectInstance", "PixelFilter",
            "ReverseOrientation", "Rotate", "Shape", "Sampler", "Scale", "TransformBegin", "TransformEnd", "TransformTimes", "Transform", "Translate",
            "WorldBegin", "WorldEnd", "Texture"]
_NUMBER = r"[+-]?(?:\d+(?:[.]\d*)?(?:[eE][+-]?\d+)?|[.]\d+(?:[eE][+-]?\d+)?)"
_TOKEN = re.compile(r"(?P<num>%s)|(?P<str>\"[^\"]*\")|(?P<kw>%s)|(?P<lb>\[)|(?P<rb>\])" % (_NUMBER, "|".join(KEYWORDS)))
_SKIP = re.compile(r"(?:[ \t\r\f\v\n]+|#[^\n]*(?:\n|$))*")
_NUM_ARRAY = re.compile(r"\[([^\]\"#]*)\]")
_NUM_RUN = re.compile(r"\s*(?:%s\s+)*%s\s*" % (_NUMBER, _NUMBER))


def parse_f32(tokens):
    """Decimal strings -> f32, correctly rounded like Rust's `parse::<f32>()`.

    numpy parses to f64 first; rounding that again to f32 is wrong only when the f64 lies within one f64-ulp of
    the midpoint of two neighbouring f32 values, so those (rare) tokens are decided in exact rational arithmetic.
    """
    d = np.array(tokens, dtype=np.float64)
    out = d.astype(f32)
    low = d.view(np.uint64) & np.uint64((1 << 29) - 1)
    half = np.uint64(1 << 28)
    sus = np.nonzero((low >= half - np.uint64(1)) & (low <= half + np.uint64(1)) & np.isfinite(d) & (np.abs(d) > 1.2e-38))[0]
    for i in sus:
        exact = Fraction(tokens[i] if isinstance(tokens[i], str) else repr(tokens[i]))
        dn = np.nextafter(out[i], f32(-np.inf))
        up = np.nextafter(out[i], f32(np.inf))
        best = out[i]
        for c in (dn, up):
            if not np.isfinite(c):
                continue
            ec, eb = abs(Fraction(float(c)) - exact), abs(Fraction(float(best)) - exact)
            if ec < eb or (ec == eb and (int(np.array(c, f32).view(np.uint32)) & 1) == 0):
                best = c
        out[i] = best
    return out


class Lexer:
    def __init__(self, text, name="<string>"):
        self.text, self.pos, self.name = text, 0, name
        self.peeked = None

    def _line(self):
        return self.text.count("\n", 0, self.pos) + 1

    def error(self, msg):
        return B200Error(f'error parsing file "{self.name}": line {self._line()}: {msg}')

    def next(self):
        if self.peeked is not None:
            t, self.peeked = self.peeked, None
            return t
        self.pos = _SKIP.match(self.text, self.pos).end()
        if self.pos >= len(self.text):
            return None
        if self.text[self.pos] == "[":  # fast path: a bracketed run of whitespace-separated numbers becomes ONE token (large meshes)
            m = _NUM_ARRAY.match(self.text, self.pos)
            if m and _NUM_RUN.fullmatch(m.group(1)):
                self.pos = m.end()
                return ("numarray", parse_f32(m.group(1).split()))
        m = _TOKEN.match(self.text, self.pos)
        if not m:
            raise self.error(f"unexpected text {self.text[self.pos:self.pos + 24]!r}")
        self.pos = m.end()
        kind = m.lastgroup
        if kind == "num":
            return ("num", m.group())
        if kind == "str":
            return ("str", m.group()[1:-1])
        if kind == "kw":
            return ("kw", m.group())
        return (kind, None)

    def peek(self):
        if self.peeked is None:
            self.peeked = self.next()
        return self.peeked


class Parser:
    """The grammar of commands.lalrpop, recursive descent."""

    def __init__(self, text, name="<string>"):
        self.lx = Lexer(text, name)

    def _expect(self, kind, what):
        t = self.lx.next()
        if t is None or t[0] != kind:
            raise self.lx.error(f"expected {what}, found {t}")
        return t[1]

    def _numbers(self, n, what):  # n bare Number tokens (LookAt, Scale, ...)
        toks = [self._expect("num", f"a number for {what}") for _ in range(n)]
        return [float(v) for v in parse_f32(toks)]

    def _bracketed(self):
        """`[` already consumed: Number+ `]` or Str+ `]`."""
        nums, strs = [], []
        while True:
            u = self.lx.next()
            if u is None:
                raise self.lx.error("unterminated [")
            if u[0] == "rb":
                break
            if u[0] == "num":
                nums.append(u[1])
            elif u[0] == "str":
                strs.append(u[1])
            else:
                raise self.lx.error(f"unexpected {u} inside [ ]")
        if (nums and strs) or not (nums or strs):
            raise self.lx.error("an array holds one or more numbers, or one or more strings")
        return parse_f32(nums) if nums else strs

    def _array(self, numbers_only=False):  # Array: Floats | Strings;  Floats: Number | [ Number+ ];  Strings: Str | [ Str+ ]
        t = self.lx.next()
        if t is None:
            raise self.lx.error("expected a value, found end of file")
        if t[0] == "num":
            v = parse_f32([t[1]])
        elif t[0] == "numarray":
            v = t[1]
        elif t[0] == "str":
            v = [t[1]]
        elif t[0] == "lb":
            v = self._bracketed()
        else:
            raise self.lx.error(f"expected a value, found {t}")
        if numbers_only and isinstance(v, list):
            raise self.lx.error("expected numbers")
        return v

    def _floats(self):
        return self._array(numbers_only=True)

    def _params(self):  # Params: (Str Array)*
        ps = PS.ParamSet()
        while True:
            t = self.lx.peek()
            if t is None or t[0] != "str":
                return ps
            self.lx.next()
            try:
                ps.add(t[1], self._array())
            except ValueError as e:
                raise self.lx.error(str(e))

    def commands(self):
        """Yields (directive, args...) in file order; `Commands = Cmd+`."""
        n = 0
        while True:
            t = self.lx.next()
            if t is None:
                if n == 0:
                    raise self.lx.error("no directives")
                return
            if t[0] != "kw":
                raise self.lx.error(f"expected a directive, found {t}")
            k = t[1]
            n += 1
            if k in ("AttributeBegin", "AttributeEnd", "TransformBegin", "TransformEnd", "ObjectEnd", "WorldBegin", "WorldEnd", "ReverseOrientation",
                     "Identity"):
                yield (k,)
            elif k == "ActiveTransform":
                u = self.lx.next()
                if u is None or u[0] != "kw" or u[1] not in ("All", "StartTime", "EndTime"):
                    raise self.lx.error("ActiveTransform takes All, StartTime or EndTime")
                yield (k, u[1])
            elif k in ("Accelerator", "Camera", "Film", "Integrator", "AreaLightSource", "LightSource", "Material", "MakeNamedMaterial", "MakeNamedMedium",
                       "Sampler", "Shape", "PixelFilter"):
                yield (k, self._expect("str", f"a name after {k}"), self._params())
            elif k in ("ObjectBegin", "ObjectInstance", "CoordinateSystem", "CoordSysTransform", "Include", "NamedMaterial"):
                yield (k, self._expect("str", f"a name after {k}"))
            elif k == "MediumInterface":
                yield (k, self._expect("str", "inside medium"), self._expect("str", "outside medium"))
            elif k == "LookAt":
                yield (k, self._numbers(9, k))
            elif k in ("Scale", "Translate"):
                yield (k, self._numbers(3, k))
            elif k == "Rotate":
                yield (k, self._numbers(4, k))
            elif k == "TransformTimes":
                yield (k, self._numbers(2, k))
            elif k in ("ConcatTransform", "Transform"):
                yield (k, self._floats())
            elif k == "Texture":
                name, ty, texname = (self._expect("str", "Texture name / type / class") for _ in range(3))
                yield (k, name, ty, texname, self._params())
            else:
                raise self.lx.error(f"{k} is not a directive")


def parse_commands(text, name="<string>"):
    return list(Parser(text, name).commands())


def run_commands(commands, api):
    """pbrtparser::parse, pbrtparser.rs:38-91."""
    for c in commands:
        k = c[0]
        if k == "ActiveTransform":
            {"All": api.active_transform_all, "StartTime": api.active_transform_starttime, "EndTime": api.active_transform_endtime}[c[1]]()
        elif k == "AttributeBegin":
            api.attribute_begin()
        elif k == "AttributeEnd":
            api.attribute_end()
        elif k == "TransformBegin":
            api.transform_begin()
        elif k == "TransformEnd":
            api.transform_end()
        elif k == "ObjectEnd":
            api.object_end()
        elif k == "WorldBegin":
            api.world_begin()
        elif k == "WorldEnd":
            api.world_end()
        elif k == "ReverseOrientation":
            api.reverse_orientation()
        elif k == "Identity":
            api.identity()
        elif k == "Accelerator":
            api.accelerator(c[1], c[2])
        elif k == "ObjectBegin":
            api.object_begin(c[1])
        elif k == "ObjectInstance":
            api.object_instance(c[1])
        elif k == "LookAt":
            api.lookat(*c[1])
        elif k == "CoordinateSystem":
            api.coordinate_system(c[1])
        elif k == "CoordSysTransform":
            api.coord_sys_transform(c[1])
        elif k == "Camera":
            api.camera(c[1], c[2])
        elif k == "Film":
            api.film(c[1], c[2])
        elif k == "Integrator":
            api.integrator(c[1], c[2])
        elif k == "AreaLightSource":
            api.area_lightsource(c[1], c[2])
        elif k == "LightSource":
            api.light_source(c[1], c[2])
        elif k == "Material":
            api.material(c[1], c[2])
        elif k == "MakeNamedMaterial":
            api.make_named_material(c[1], c[2])
        elif k == "MakeNamedMedium":
            api.make_named_medium(c[1], c[2])
        elif k == "NamedMaterial":
            api.named_material(c[1])
        elif k == "Sampler":
            api.sampler(c[1], c[2])
        elif k == "Shape":
            api.shape(c[1], c[2])
        elif k == "PixelFilter":
            api.pixel_filter(c[1], c[2])
        elif k == "Scale":
            api.scale(*c[1])
        elif k == "Rotate":
            api.rotate(*c[1])
        elif k == "Translate":
            api.translate(*c[1])
        elif k == "TransformTimes":
            api.transform_times(*c[1])
        elif k == "Texture":
            api.texture(c[1], c[2], c[3], c[4])
        elif k == "ConcatTransform":
            api.concat_transform(c[1])
        elif k == "Transform":
            api.transform(c[1])
        elif k == "Include":
            api.include(c[1])
        elif k == "MediumInterface":
            api.medium_interface("" if c[1] == '""' else c[1], "" if c[2] == '""' else c[2])
        else:
            raise B200Error(f"unhandled directive {k}")


def parse_file(path, api):
    """pbrtparser::parse: the whole file is parsed before its first directive runs (pbrtparser.rs:38-40)."""
    with open(path, "r") as f:
        text = f.read()
    run_commands(parse_commands(text, str(path)), api)


def pbrt_parse(path, **options):
    """pbrtparser::pbrt_parse (pbrtparser.rs:30-36): returns the API after the file ran; `api.jobs` holds one
    RenderJob per WorldEnd (the reference renders right there, api.rs:1740-1747; here rendering needs a device,
    so WorldEnd flattens the scene and the caller launches `job.render(device)`)."""
    from .api import API

    PS.set_search_directory(os.path.dirname(os.path.abspath(str(path))))
    api = API(**options)
    parse_file(path, api)
    return api


def pbrt_parse_string(text, search_dir=None, **options):
    from .api import API

    PS.set_search_directory(search_dir)
    api = API(**options)
    run_commands(parse_commands(text), api)
    return api
