"""Host-side BVH build of the product library (pbrt-rust_b200/csrc/host_bvh.cpp: BVHAccel::new, bvh.rs:145-375,
662-693) against the oracle's restatement: node arrays and `ordered_prims` must be identical, including the
reference's quirks (right subtree built first, nprims<=2 median split, degenerate-centroid leaves)."""
import numpy as np
import pytest


def _bounds(rng, n, degenerate=False):
    c = rng.uniform(-10, 10, size=(n, 3)).astype(np.float32)
    if degenerate:
        c[:] = c[0]
    e = rng.uniform(0.01, 0.5, size=(n, 3)).astype(np.float32)
    return np.concatenate([c - e, c + e], axis=1).astype(np.float32)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 17, 255, 1000, 20000])
@pytest.mark.parametrize("method", ["sah", "middle", "equal"])
def test_bvh_build_equals_oracle(pkg, oracle, n, method):
    pb = _bounds(np.random.RandomState(n), n)
    nodes, ordered = pkg.bvh_build(pb, 4, method)
    onodes, oordered = oracle.bvh_build(pb, 4, method)
    assert np.array_equal(ordered, oordered)
    assert nodes.tobytes() == onodes.tobytes()
    assert sorted(ordered.tolist()) == list(range(n))


def test_bvh_structure_invariants(pkg):
    pb = _bounds(np.random.RandomState(3), 5000)
    nodes, ordered = pkg.bvh_build(pb, 4, "sah")
    # depth-first layout: first child at i+1, second child at `offset` (bvh.rs:662-693); leaves partition ordered_prims
    covered = np.zeros(len(ordered), bool)
    stack = [0]
    while stack:
        i = stack.pop()
        nd = nodes[i]
        if nd["n_prims"] > 0:
            sl = slice(int(nd["offset"]), int(nd["offset"]) + int(nd["n_prims"]))
            assert not covered[sl].any()
            covered[sl] = True
            b = pb[ordered[sl]]
            assert np.all(b[:, :3] >= nd["bounds"][:3]) and np.all(b[:, 3:] <= nd["bounds"][3:])
        else:
            assert nd["axis"] in (0, 1, 2) and nd["offset"] > i + 1
            stack += [i + 1, int(nd["offset"])]
    assert covered.all()


def test_degenerate_centroids_make_a_leaf(pkg, oracle):
    pb = _bounds(np.random.RandomState(1), 40, degenerate=True)
    pb[:, :3] = pb[0, :3]; pb[:, 3:] = pb[0, 3:]
    nodes, ordered = pkg.bvh_build(pb, 4, "sah")
    onodes, oordered = oracle.bvh_build(pb, 4, "sah")
    assert nodes.tobytes() == onodes.tobytes() and np.array_equal(ordered, oordered)
    assert len(nodes) == 1 and nodes[0]["n_prims"] == 40  # bvh.rs:238-247


def test_empty_scene_tables(pkg):
    flat = pkg.SceneBuilder().world_end()
    assert len(flat.nodes) == 0 and len(flat.prims) == 0
