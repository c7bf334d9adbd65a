"""The oracle against its own frozen closest-hit answers (tests/golden/closest_hit_vectors.npz, made by
tests/golden/make_closest_hit_vectors.py): the checker of the GPU parity tests must not drift.  The scene functions
(host/scenes.cpp mirroring main.rs:192-795), the lowering and the oracle's `BVHNode::hit` (bvh.rs:25-50) all take part."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "closest_hit_vectors.npz")
KW = {"stress": dict(stress_spheres=100_000)}


def load_golden():
    z = np.load(GOLDEN)
    return z, [str(s) for s in z["scenes"]], int(z["seed"])


@pytest.mark.parametrize("name", load_golden()[1])
def test_oracle_reproduces_golden_closest_hits(rt, oracle, name):
    z, _, seed = load_golden()
    hs = rt.api.HostScene(name, seed=1, **KW.get(name, {}))
    osc = oracle.OracleScene(hs.desc)
    prim, t, normal, ff, uv, amb = osc.trace_closest(z[name + "/rays"], seed=seed)
    assert np.array_equal(amb.astype(np.uint8), z[name + "/ambiguous"])
    assert np.array_equal(prim, z[name + "/prim"])  # ties included: the oracle's visiting order is part of what is frozen
    hit = prim >= 0
    assert hit.sum() >= 0.05 * len(prim)
    gt = z[name + "/t"]
    assert np.all(np.abs(t[hit] - gt[hit]) <= 1e-12 * np.abs(gt[hit]))  # (compiler FMA contraction may move the last bits)
    assert np.isinf(t[~hit]).all() and np.isinf(gt[~hit]).all()
    assert np.abs(normal[hit] - z[name + "/normal"][hit]).max() <= 1e-6
    assert np.array_equal(ff[hit].astype(np.uint8), z[name + "/front_face"][hit])
    assert np.abs(uv[hit] - z[name + "/uv"][hit]).max() <= 1e-6
    osc.close()


def test_golden_vectors_cover_every_primitive_kind(rt):
    """The frozen ray sets reach spheres, moving spheres, all three rectangle kinds, box sides under wrappers and media."""
    api = rt.api
    z, names, _ = load_golden()
    kinds = set()
    for name in names:
        hs = api.HostScene(name, seed=1, **KW.get(name, {}))
        prims = api.lower_prims(hs.desc)
        ids = z[name + "/prim"]
        kinds |= {prims[i].kind for i in set(ids[ids >= 0].tolist())}
    want = {api.NODE_SPHERE, api.NODE_MOVING_SPHERE, api.NODE_XY_RECT, api.NODE_XZ_RECT, api.NODE_YZ_RECT, api.NODE_CONSTANT_MEDIUM}
    assert want <= kinds, f"missing primitive kinds: {want - kinds}"
