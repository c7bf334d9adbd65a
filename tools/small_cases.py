"""Small renders (64 x 64 and below, wave capacity 4096) and closest-hit batches through every wave-kernel variant:
flat / BVH lockstep / BVH persistent, with and without media and rich textures.  A seconds-long smoke of the whole
kernel matrix, also the right size for a run under a memory checker where one is available.

    python tools/small_cases.py
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
                                    ("final_scene", 32, 2, 0), ("final_scene", 32, 1, api.FLAG_BVH_PERSISTENT)):
        hs = api.HostScene(name, seed=1)
        scene = api.Scene(ctx, hs.desc)
        img, _, st = scene.render(hs.camera(), hs.params(spp=spp, width=width, flags=flags, pool_paths=4096))
        rays = np.zeros(2048, dtype=api.RAY_DTYPE)
        rays["origin"] = np.array(list(hs.settings.look_from)) + rng.normal(size=(2048, 3))
        rays["direction"] = np.array(list(hs.settings.look_at)) - rays["origin"] + rng.normal(size=(2048, 3)) * 50.0
        prim = scene.trace_closest(rays)[0]
        print(f"{name:20s} {img.shape} paths {st.paths} rays {st.rays} waves {st.waves} hits {(prim >= 0).sum()}", flush=True)
        scene.close()
    ctx.close()


if __name__ == "__main__":
    main()
