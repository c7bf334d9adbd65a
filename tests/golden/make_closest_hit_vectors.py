"""tests/golden/closest_hit_vectors.npz: closest-hit known answers of the CPU oracle on a small fixed ray set per scene.

    python tests/golden/make_closest_hit_vectors.py        (CPU only; needs the built oracle and host library)

The reference ships no test vectors for `BVHNode::hit`, and the Rust crate cannot run here, so these
are the ORACLE's answers (oracle/oracle.cpp: bvh.rs:25-50 and the `Hittable::hit` impls it calls), frozen: the CPU suite
checks that the oracle still gives them (a guard against drift of the checker), the GPU suite checks `rt1w_trace_closest`
against them with the bars of tests/common.py: check_trace_parity.  Rays: tests/common.py: make_ray_set (half uniform in
the scene box, a quarter camera rays, a quarter secondary rays), stored with the answers so that the fixture does not
depend on numpy's generators.  Medium hits (cornel_smoke, final_scene) replay the Philox draws of `seed` (kSeed).
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SEED = 0x5EED
N = 1024
# name, HostScene keywords, max_extent of the uniform rays (scenes with a huge ground / fog sphere), aspect
CASES = [("cornel_box", {}, None), ("cornel_smoke", {}, None), ("simple_light", {}, 30.0), ("random_scene", {}, 15.0),
         ("final_scene", {}, 700.0), ("stress", dict(stress_spheres=100_000), None)]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "closest_hit_vectors.npz")


def main():
    api = importlib.import_module("raytracing-1w_b200").api
    import oracle_binding
    from common import make_ray_set

    out = {"seed": np.uint64(SEED), "scenes": np.array([c[0] for c in CASES])}
    for name, kw, extent in CASES:
        hs = api.HostScene(name, seed=1, **kw)
        osc = oracle_binding.OracleScene(hs.desc)
        rays = make_ray_set(api, hs, osc, api.lower_prims(hs.desc), N, extent)
        prim, t, normal, ff, uv, amb = osc.trace_closest(rays, seed=SEED)
        out[name + "/rays"] = rays
        out[name + "/prim"] = prim.astype(np.int32)
        out[name + "/t"] = t.astype(np.float64)
        out[name + "/normal"] = normal.astype(np.float32)
        out[name + "/front_face"] = ff.astype(np.uint8)
        out[name + "/uv"] = uv.astype(np.float32)
        out[name + "/ambiguous"] = amb.astype(np.uint8)
        print(f"{name}: {N} rays, {(prim >= 0).mean():.3f} hit, {(amb != 0).mean():.4f} ambiguous, {len(set(prim.tolist()))} distinct primitives")
        osc.close()
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
