"""Scenes the reference's `main.rs` never builds but its types allow (built through the POD description, not host/scenes.cpp):
the corners of the lowering and of the kernel selection that the eight shipped scenes do not reach.

  * many AABoxes and no wrapper: more than 32 rectangles that merge to fewer than 32 box primitives (must take the BVH path
    with P_BOX leaves - the flat scan compiles P_BOX out);
  * image textures on rectangles and box sides (aarect.rs:60-61 sets u, v; texture.rs:67-89 reads them), also under wrappers;
  * several Perlin tables (static + dynamic shared memory beyond 48 KB needs the opt-in; too many tables fall back to global);
  * wrappers ABOVE a ConstantMedium (hittable.rs:221-229,269-277,290-294 rewrite the medium's literal record).
Closest hits against the oracle with the bars of test_gpu_trace_parity.py, images with those of test_gpu_render_parity.py.
"""
import ctypes as C

import numpy as np
import pytest

from common import check_render_parity, check_trace_parity, make_ray_set

pytestmark = pytest.mark.gpu


class AdHoc:
    """What make_ray_set / the render helpers need of a host scene, for a description built in Python."""

    class _Settings:
        pass

    def __init__(self, api, builder, look_from, look_at, vfov, background, aspect=1.0):
        self.api, self.b = api, builder
        self.desc = builder.desc()
        self.settings = self._Settings()
        self.settings.look_at = look_at
        self.look_from, self.vfov, self.aspect, self.background = look_from, vfov, aspect, background

    def camera(self):
        return self.api.camera_new(self.look_from, self.settings.look_at, (0.0, 1.0, 0.0), self.vfov, self.aspect, 0.0, 10.0, 0.0, 1.0)

    def params(self, width, spp, sample_begin=0, flags=0, stat_clamp=0.0, seed=0, max_depth=50):
        p = self.api.RenderParams()
        p.width, p.height = width, int(width / self.aspect)
        p.sample_begin, p.sample_end, p.max_depth, p.flags, p.seed = sample_begin, spp, max_depth, flags, seed
        p.background[:] = list(self.background)
        p.stat_clamp = stat_clamp
        return p

    def make_params(self, width):
        return lambda spp, begin, flags, clamp, seed: self.params(width, spp, begin, flags, clamp, seed)


def _random_image(rng, h=8, w=16):
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


def _perlin_tables(rng):
    v = rng.uniform(-1.0, 1.0, (256, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v, rng.permutation(256), rng.permutation(256), rng.permutation(256)


def _room(b, size=555.0, light=(213.0, 343.0, 227.0, 332.0)):
    """The five walls and the ceiling light of main.rs:451-494 (colours aside)."""
    red, white, green = (b.lambertian(b.solid(*c)) for c in ((0.65, 0.05, 0.05), (0.73, 0.73, 0.73), (0.12, 0.45, 0.15)))
    lamp = b.diffuse_light(b.solid(15.0, 15.0, 15.0))
    kids = [b.yz_rect(0, size, 0, size, size, green), b.yz_rect(0, size, 0, size, 0, red),
            b.flip_face(b.xz_rect(light[0], light[1], light[2], light[3], size - 1.0, lamp)),
            b.xz_rect(0, size, 0, size, 0, white), b.xz_rect(0, size, 0, size, size, white), b.xy_rect(0, size, 0, size, size, white)]
    return kids, white


def boxes_scene(api):
    """6 walls/light + 6 AABoxes, no wrapper frames: 42 rectangles -> 12 device primitives (6 P_BOX leaves in a BVH)."""
    rng = np.random.Generator(np.random.Philox(11))
    b = api.DescBuilder()
    kids, white = _room(b)
    mats = [b.lambertian(b.image(_random_image(rng))), b.metal((0.8, 0.85, 0.88), 0.0), b.dielectric(1.5), white,
            b.lambertian(b.checker(b.solid(0.2, 0.3, 0.1), b.solid(0.9, 0.9, 0.9))), b.metal((0.7, 0.6, 0.5), 0.3)]
    for k, m in enumerate(mats):
        x0, z0 = 60.0 + 165.0 * (k % 3), 80.0 + 220.0 * (k // 3)
        kids.append(b.aabox((x0, 0.0 if k % 2 else 40.0, z0), (x0 + 110.0, 120.0 + 45.0 * k, z0 + 130.0), m))
    b.set_world(b.bvh(kids))
    b.set_lights([b.xz_rect(213.0, 343.0, 227.0, 332.0, 554.0, b.null_material())])
    return AdHoc(api, b, (278.0, 278.0, -800.0), (278.0, 278.0, 0.0), 40.0, (0.0, 0.0, 0.0))


def image_rects_scene(api):
    """Image textures on a wall, on a rotated + translated box and on a rectangle under a Translate: a flat-scan scene."""
    rng = np.random.Generator(np.random.Philox(12))
    b = api.DescBuilder()
    kids, white = _room(b)
    kids[-1] = b.xy_rect(0, 555.0, 0, 555.0, 555.0, b.lambertian(b.image(_random_image(rng, 16, 16))))  # back wall
    box = b.aabox((0.0, 0.0, 0.0), (165.0, 330.0, 165.0), b.lambertian(b.image(_random_image(rng))))
    kids.append(b.translate(b.rotate_y(box, 15.0), (265.0, 0.0, 295.0)))
    kids.append(b.translate(b.yz_rect(0.0, 200.0, 0.0, 150.0, 0.0, b.lambertian(b.image(_random_image(rng, 4, 4)))), (120.0, 30.0, 60.0)))
    b.set_world(b.bvh(kids))
    b.set_lights([b.xz_rect(213.0, 343.0, 227.0, 332.0, 554.0, b.null_material())])
    return AdHoc(api, b, (278.0, 278.0, -800.0), (278.0, 278.0, 0.0), 40.0, (0.0, 0.0, 0.0))


def perlin_scene(api, n_tables, n_extra_spheres=0):
    """Spheres with a NoiseTexture each (every one its own Perlin<256>, perlin.rs:14-44) under a sky."""
    rng = np.random.Generator(np.random.Philox(13))
    b = api.DescBuilder()
    kids = [b.sphere((0.0, -1000.0, 0.0), 1000.0, b.lambertian(b.noise(4.0, *_perlin_tables(rng))))]
    for k in range(1, n_tables):
        kids.append(b.sphere((2.5 * (k - n_tables / 2.0), 1.0, 1.5 * (k % 2)), 1.0, b.lambertian(b.noise(2.0 + k, *_perlin_tables(rng)))))
    grey = b.lambertian(b.solid(0.5, 0.5, 0.5))
    for k in range(n_extra_spheres):
        kids.append(b.sphere((rng.uniform(-8, 8), 0.2, rng.uniform(2, 8)), 0.2, grey))
    b.set_world(b.bvh(kids))
    return AdHoc(api, b, (13.0, 2.0, 3.0), (0.0, 0.0, 0.0), 30.0, (0.7, 0.8, 1.0), aspect=1.5)


def wrapped_media_scene(api):
    """ConstantMedium under Translate / RotateY / FlipFace wrappers, next to plain surfaces."""
    b = api.DescBuilder()
    kids, white = _room(b)
    none = b.null_material()  # a boundary's own material is never read (constant_medium.rs:58-72)
    fog = b.constant_medium(b.sphere((0.0, 0.0, 0.0), 80.0, none), 0.02, b.solid(0.9, 0.9, 0.9))
    kids.append(b.translate(b.rotate_y(fog, 30.0), (190.0, 90.0, 190.0)))
    smoke = b.constant_medium(b.translate(b.rotate_y(b.aabox((0, 0, 0), (165, 330, 165), none), 15.0), (265, 0, 295)), 0.01, b.solid(0.1, 0.1, 0.1))
    kids.append(b.flip_face(smoke))
    kids.append(b.rotate_y(b.translate(b.constant_medium(b.aabox((0, 0, 0), (100, 100, 100), none), 0.05, b.solid(0.5, 0.7, 0.9)), (60, 300, 100)), -10.0))
    b.set_world(b.bvh(kids))
    b.set_lights([b.xz_rect(213.0, 343.0, 227.0, 332.0, 554.0, b.null_material())])
    return AdHoc(api, b, (278.0, 278.0, -800.0), (278.0, 278.0, 0.0), 40.0, (0.0, 0.0, 0.0))


def enclosed_scene(api):
    """Two primitives whose boxes contain everything else: a sky sphere (kept out of the tree and tested once per ray after
    the traversal, kernels.cuh: hit_globals carries the sphere tests only) and, inside it, a closed AABox room around 40
    spheres and boxes (an AABox is NOT taken out of the tree: it stays a P_BOX leaf).  The camera sits inside the room and
    glass panes let rays out to the sky sphere."""
    rng = np.random.Generator(np.random.Philox(17))
    b = api.DescBuilder()
    kids = [b.flip_face(b.sphere((0.0, 0.0, 0.0), 400.0, b.diffuse_light(b.solid(0.6, 0.7, 1.0)))),  # the sky, emitting inward (material.rs:168-181)
            b.aabox((-100.0, -100.0, -100.0), (100.0, 100.0, 100.0), b.dielectric(1.5))]      # the room: glass all around
    grey, gold = b.lambertian(b.solid(0.5, 0.5, 0.5)), b.metal((0.8, 0.6, 0.2), 0.1)
    for k in range(40):
        c = rng.uniform(-80.0, 80.0, 3)
        if k % 4 == 0:
            kids.append(b.aabox(tuple(c - rng.uniform(3.0, 9.0, 3)), tuple(c + rng.uniform(3.0, 9.0, 3)), gold if k % 8 else grey))
        else:
            kids.append(b.sphere(tuple(c), float(rng.uniform(3.0, 10.0)), grey if k % 3 else gold))
    b.set_world(b.bvh(kids))
    return AdHoc(api, b, (0.0, 20.0, -90.0), (0.0, 0.0, 0.0), 60.0, (0.0, 0.0, 0.0))


SCENES = {
    "enclosed": (enclosed_scene, {}, 1 << 16, 120.0),
    "boxes": (boxes_scene, {}, 1 << 16, None),
    "image_rects": (image_rects_scene, {}, 1 << 16, None),
    "perlin3_flat": (perlin_scene, dict(n_tables=3), 1 << 15, 12.0),
    "perlin8_bvh": (perlin_scene, dict(n_tables=8, n_extra_spheres=40), 1 << 15, 12.0),
    "perlin12_global": (perlin_scene, dict(n_tables=12, n_extra_spheres=40), 1 << 15, 12.0),
    "wrapped_media": (wrapped_media_scene, {}, 1 << 16, None),
}


@pytest.mark.parametrize("name", list(SCENES))
def test_adhoc_closest_hit_matches_oracle(rt, oracle, gpu_ctx, name):
    api = rt.api
    make, kw, n, extent = SCENES[name]
    hs = make(api, **kw)
    osc = oracle.OracleScene(hs.desc)
    gsc = api.Scene(gpu_ctx, hs.desc)
    info = gsc.info()
    if name == "enclosed":  # the sky sphere is tested outside the tree, the room box inside it
        assert info.n_global_prims == 1 and info.n_bvh_nodes > 0
    if name == "boxes":  # the six boxes travel as one primitive each: 6 + 6 leaves -> a BVH, not the scan
        assert info.n_prims == 42 and info.n_bvh_nodes >= 2 * 12 - 1
    rays = make_ray_set(api, hs, osc, gsc.prims(), n, extent)
    check_trace_parity(gsc, osc, rays, label=name)
    gsc.close()


@pytest.mark.parametrize("name,width,spp", [("enclosed", 64, 128), ("boxes", 64, 256), ("image_rects", 64, 256), ("perlin3_flat", 96, 64),
                                            ("perlin8_bvh", 96, 64), ("perlin12_global", 96, 64), ("wrapped_media", 48, 256)])
def test_adhoc_render_matches_oracle(rt, oracle, gpu_ctx, name, width, spp):
    api = rt.api
    make, kw, _, _ = SCENES[name]
    hs = make(api, **kw)
    osc = oracle.OracleScene(hs.desc)
    gsc = api.Scene(gpu_ctx, hs.desc)
    check_render_parity(api, gsc, osc, hs.camera(), hs.make_params(width), spp)
    gsc.close()
