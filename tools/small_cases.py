"""Small renders (64 x 64 and below, wave capacity 4096) and closest-hit batches through every wave-kernel variant:
flat / BVH lockstep / BVH persistent, binary / 8-wide tree, with and without media and rich textures, plus the parity hooks
(rt1w_eval_*).  A seconds-long smoke of the whole kernel matrix, also the right size for a run under a memory checker:

    python tools/small_cases.py
    compute-sanitizer --tool memcheck python tools/small_cases.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    api = importlib.import_module("raytracing-1w_b200").api
    ctx = api.Context(0)
    rng = np.random.default_rng(1)
    for name, width, spp, flags in (("cornel_box", 64, 4, 0), ("cornel_smoke", 48, 2, 0), ("two_perlin_spheres", 48, 2, 0), ("earth", 48, 2, 0),
                                    ("simple_light", 48, 2, 0), ("random_scene", 48, 2, 0), ("one_weekend", 48, 2, 0),
                                    ("final_scene", 32, 2, 0), ("final_scene", 32, 1, api.FLAG_BVH_PERSISTENT),
                                    ("final_scene", 32, 1, api.FLAG_BVH_WIDE), ("final_scene", 32, 1, api.FLAG_BVH_WIDE | api.FLAG_BVH_PERSISTENT),
                                    ("random_scene", 32, 1, api.FLAG_BVH_WIDE | api.FLAG_BVH_PERSISTENT), ("stress", 64, 1, 0),
                                    ("stress", 64, 1, api.FLAG_BVH_BINARY | api.FLAG_BVH_LOCKSTEP)):
        hs = api.HostScene(name, seed=1, **({"stress_spheres": 40_000} if name == "stress" else {}))
        scene = api.Scene(ctx, hs.desc)
        img, _, st = scene.render(hs.camera(), hs.params(spp=spp, width=width, flags=flags, pool_paths=4096))
        rays = np.zeros(2048, dtype=api.RAY_DTYPE)
        rays["origin"] = np.array(list(hs.settings.look_from)) + rng.normal(size=(2048, 3))
        rays["direction"] = np.array(list(hs.settings.look_at)) - rays["origin"] + rng.normal(size=(2048, 3)) * 50.0
        prim = scene.trace_closest(rays)[0]
        scene.eval_scatter(rays[:512], seed=3)
        if scene.info().n_lights:
            scene.eval_light_pdf(rays["origin"][:256].astype(np.float64), rays["direction"][:256])
        for tex in range(hs.desc.contents.n_textures):
            scene.eval_texture(tex, rays["origin"][:64].astype(np.float64), np.abs(rays["direction"][:64, :2]) % 1.0)
        for tab in range(hs.desc.contents.n_perlins):
            scene.eval_perlin(tab, rays["origin"][:64].astype(np.float64), turb_depth=7)
        print(f"{name:20s} {img.shape} paths {st.paths} rays {st.rays} waves {st.waves} hits {(prim >= 0).sum()}", flush=True)
        scene.close()
    ctx.eval_dielectric(np.array([[0.6, -0.8, 0.0]] * 8), np.array([[0.0, 1.0, 0.0]] * 8), np.full(8, 1.5))
    ctx.close()


if __name__ == "__main__":
    main()
