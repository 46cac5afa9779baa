"""Property tests of the scene-file front end (hypothesis): whatever f32 numbers, names and directive sequences a scene holds,
writing it with scenefile.py's number formatting and reading it back with the lexer / grammar gives the same values bit for bit,
and the API's transform stack composes them in the order the directives appear."""
import importlib

import numpy as np
from hypothesis import given, settings, strategies as st

pkg = importlib.import_module("pbrt-rust_b200")
PP, SF, H = pkg.pbrtparser, importlib.import_module("pbrt-rust_b200.scenefile"), pkg.host
f32 = np.float32

finite_f32 = st.floats(width=32, allow_nan=False, allow_infinity=False)
names = st.text(alphabet="abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_-.", min_size=1, max_size=12)


@settings(max_examples=200, deadline=None)
@given(st.lists(finite_f32, min_size=1, max_size=40))
def test_numbers_round_trip_through_text(values):
    x = np.array(values, f32)
    text = 'Shape "sphere" "float v" [' + " ".join(SF._num(v) for v in x) + "]"
    ps = PP.parse_commands(text)[0][2]
    got = ps.find_float("v")
    assert got.dtype == f32 and got.tobytes() == x.tobytes()
    # the same numbers as bare tokens (no bracketed fast path): one token at a time through the same rounding
    assert PP.parse_f32([SF._num(v) for v in x]).tobytes() == x.tobytes()


@settings(max_examples=100, deadline=None)
@given(st.lists(st.tuples(st.sampled_from(["float", "integer", "string", "bool", "rgb", "point", "vector", "normal", "point2"]), names), min_size=1, max_size=8,
                unique_by=lambda t: t[1]), st.randoms(use_true_random=False))
def test_parameter_lists_round_trip(decls, rnd):
    parts, want = [], {}
    for ty, name in decls:
        if ty == "string":
            v = [name[::-1] or "x"]
            parts.append(f'"string {name}" "{v[0]}"')
        elif ty == "bool":
            v = [rnd.random() < 0.5]
            parts.append(f'"bool {name}" ["{"true" if v[0] else "false"}"]')
        elif ty == "integer":
            v = np.array([rnd.randint(-1000, 1000) for _ in range(rnd.randint(1, 6))], np.int64)
            parts.append(f'"integer {name}" [' + " ".join(str(int(i)) for i in v) + "]")
        else:
            k = {"float": 1, "rgb": 3, "point": 3, "vector": 3, "normal": 3, "point2": 2}[ty]
            v = np.array([rnd.uniform(-1e3, 1e3) for _ in range(k * rnd.randint(1, 4))], f32)
            parts.append(f'"{ty} {name}" [' + " ".join(SF._num(x) for x in v) + "]")
        want[(ty, name)] = v
    ps = PP.parse_commands('Material "matte" ' + " ".join(parts))[0][2]
    bucket = {"float": "floats", "integer": "ints", "string": "strings", "bool": "bools", "rgb": "spectra", "point": "point3fs", "vector": "vector3fs",
              "normal": "normals", "point2": "point2fs"}
    for (ty, name), v in want.items():
        got = getattr(ps, bucket[ty])[name]
        assert np.array_equal(np.asarray(got).reshape(-1), np.asarray(v).reshape(-1)), (ty, name)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.sampled_from(["Translate", "Scale", "Rotate"]), st.tuples(*[st.floats(min_value=0.25, max_value=4.0, width=32)] * 4)), min_size=1, max_size=6))
def test_transform_directives_compose_in_file_order(ops):
    lines, t = ["WorldBegin"], H.Transform()
    for op, (a, b, c, d) in ops:
        if op == "Translate":
            lines.append(f"Translate {SF._num(a)} {SF._num(b)} {SF._num(c)}"); t = t * H.Transform.translate((a, b, c))
        elif op == "Scale":
            lines.append(f"Scale {SF._num(a)} {SF._num(b)} {SF._num(c)}"); t = t * H.Transform.scale(a, b, c)
        else:
            lines.append(f"Rotate {SF._num(a * 30)} {SF._num(b)} {SF._num(c)} {SF._num(d)}"); t = t * H.Transform.rotate(f32(a * 30), (b, c, d))
    lines += ['Shape "sphere"', "WorldEnd"]
    job = pkg.pbrt_parse_string("\n".join(lines)).jobs[0]
    assert job.flat.spheres[0]["object_to_world"].tobytes() == t.m.reshape(-1).astype(f32).tobytes()
    assert job.flat.spheres[0]["world_to_object"].tobytes() == t.m_inv.reshape(-1).astype(f32).tobytes()
