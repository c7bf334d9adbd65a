"""Throughput of the BASELINE.json configurations other than the headline (parity-test cases, not bench lines).

    python tools/scene_perf.py [scene[:spp[:width]] ...]      # prints one JSON line per scene
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = ["cornel_box:100", "random_scene:32", "final_scene:32", "cornel_smoke:32", "stress:8"]


def main():
    rt = importlib.import_module("raytracing-1w_b200")
    api = rt.api
    ctx = api.Context(0)
    for spec in (sys.argv[1:] or DEFAULT):
        name, _, rest = spec.partition(":")
        spp, _, width = rest.partition(":")
        spp, width = int(spp or 16), (int(width) if width else None)
        t0 = time.perf_counter()
        hs = api.HostScene(name, seed=1, **({"stress_spheres": int(os.environ.get("RT1W_STRESS_SPHERES", "1000000"))} if name == "stress" else {}))
        t1 = time.perf_counter()
        scene = api.Scene(ctx, hs.desc)
        info = scene.info()
        cam = hs.camera()
        flags = int(os.environ.get("RT1W_FLAGS", "0"))  # e.g. 16 | 4: binary tree, lockstep kernel (include/rt1w.h: RT1W_FLAG_BVH_*)
        p = hs.params(spp=spp, width=width, flags=flags)
        if os.environ.get("RT1W_NO_WARMUP"):  # one render only (ncu captures: launch k is wave k)
            img, _, st = scene.render(cam, p)
            sp = st
        else:
            scene.render(cam, hs.params(spp=1, width=width))  # warm-up (allocations, module load)
            img, _, st = scene.render(cam, p)
            pp = hs.params(spp=spp, width=width, flags=api.FLAG_PROFILE | flags)
            _, _, sp = scene.render(cam, pp)
        print(json.dumps({
            "scene": name, "image": [p.width, p.height], "spp": spp, "prims": info.n_prims, "bvh_nodes": info.n_bvh_nodes,
            "bvh_depth": info.bvh_depth, "wide_nodes": info.n_wide_nodes, "wide_depth": info.wide_depth, "wide_children": round(info.wide_children, 2), "flags": flags, "host_scene_s": round(t1 - t0, 3), "build_ms": round(info.build_ms, 1),
            "upload_ms": round(info.upload_ms, 1), "render_ms": round(st.render_ms, 2), "mpaths_s": round(st.paths / st.render_ms / 1e3, 1),
            "mrays_s": round(st.rays / st.render_ms / 1e3, 1), "rays_per_path": round(st.rays / st.paths, 3), "waves": st.waves,
            "kernel_ms": {api.KERNEL_NAMES[k]: round(sp.kernel_ms[k], 2) for k in range(7) if sp.kernel_launches[k]},
            "nan_pixels": int((img != img).any(axis=2).sum()),
            # A/B of code that must not change results: the ray count is exact (every path decision is deterministic), the image sum
            # equal up to the order of the fp32 atomic additions
            "rays": int(st.rays), "image_sum": float(np.nansum(img.astype(np.float64)))}), flush=True)
        scene.close()
    ctx.close()


if __name__ == "__main__":
    main()
