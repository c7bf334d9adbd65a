"""Lockstep vs persistent BVH wave kernel x binary vs compressed 8-wide tree on the BVH scenes (forces each with its render flags).

    python tools/bvh_modes.py [scene:spp[:width] ...]
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    api = importlib.import_module("raytracing-1w_b200").api
    ctx = api.Context(0)
    for spec in (sys.argv[1:] or ["one_weekend:64:1200", "random_scene:64:1200", "final_scene:32:800", "stress:8"]):
        name, spp, *w = spec.split(":")
        width = int(w[0]) if w else None
        hs = api.HostScene(name, seed=1, **({"stress_spheres": 1_000_000} if name == "stress" else {}))
        scene = api.Scene(ctx, hs.desc)
        cam = hs.camera()
        modes = [(f"{k} / {t}", kf | tf) for k, kf in (("lockstep", api.FLAG_BVH_LOCKSTEP), ("persistent", api.FLAG_BVH_PERSISTENT))
                 for t, tf in (("binary", api.FLAG_BVH_BINARY), ("8-wide", api.FLAG_BVH_WIDE))]
        for label, flag in modes:
            scene.render(cam, hs.params(spp=1, width=width, flags=flag))
            _, _, st = scene.render(cam, hs.params(spp=int(spp), width=width, flags=flag))
            print(f"{name:14s} {label:20s} {st.render_ms:8.2f} ms  {st.paths / st.render_ms / 1e3:8.1f} Mpaths/s  {st.rays / st.render_ms / 1e3:8.1f} Mrays/s", flush=True)
        scene.close()
    ctx.close()


if __name__ == "__main__":
    main()
