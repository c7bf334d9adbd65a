"""Shared helpers of the parity tests: the fixed synthetic ray set and image statistics."""
import numpy as np

RAY_SEED = 0xC0FFEE


def scene_bounds(prims):
    lo = np.min([list(p.bbox_min) for p in prims], axis=0)
    hi = np.max([list(p.bbox_max) for p in prims], axis=0)
    return lo, hi


def uniform_sphere(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def cosine_about(rng, normals):
    """Cosine-distributed directions about unit normals (math.rs:39-49 in an arbitrary frame)."""
    n = normals.shape[0]
    r1, r2 = rng.random(n), rng.random(n)
    phi = 2 * np.pi * r1
    local = np.stack([np.cos(phi) * np.sqrt(r2), np.sin(phi) * np.sqrt(r2), np.sqrt(1 - r2)], axis=1)
    a = np.where(np.abs(normals[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    v = np.cross(normals, a)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    u = np.cross(normals, v)
    return local[:, :1] * u + local[:, 1:2] * v + local[:, 2:3] * normals


def make_ray_set(api, host_scene, oracle_scene, prims, n, max_extent=None):
    """Half uniform rays in the inflated scene box, a quarter camera rays, a quarter secondary rays."""
    rng = np.random.Generator(np.random.Philox(RAY_SEED))
    lo, hi = scene_bounds(prims)
    if max_extent is not None:  # scenes with a huge fog/ground sphere: sample around the camera target instead
        ctr = np.array(list(host_scene.settings.look_at))
        lo, hi = np.maximum(lo, ctr - max_extent), np.minimum(hi, ctr + max_extent)
    ext = hi - lo
    lo, hi = lo - 0.1 * ext, hi + 0.1 * ext
    n_uni, n_cam = n // 2, n // 4
    n_sec = n - n_uni - n_cam
    rays = np.zeros(n, dtype=api.RAY_DTYPE)
    # uniform
    rays["origin"][:n_uni] = lo + rng.random((n_uni, 3)) * (hi - lo)
    rays["direction"][:n_uni] = uniform_sphere(rng, n_uni) * (0.5 + 19.5 * rng.random((n_uni, 1)))
    # camera rays on a regular (s, t) grid (camera.rs:67-70 with lens radius 0)
    cam = host_scene.camera()
    side = int(np.ceil(np.sqrt(n_cam)))
    s, t = np.meshgrid((np.arange(side) + 0.5) / side, (np.arange(side) + 0.5) / side)
    s, t = s.ravel()[:n_cam], t.ravel()[:n_cam]
    origin = np.array(list(cam.origin))
    llc, hor, ver = np.array(list(cam.lower_left_corner)), np.array(list(cam.horizontal)), np.array(list(cam.vertical))
    rays["origin"][n_uni:n_uni + n_cam] = origin
    rays["direction"][n_uni:n_uni + n_cam] = llc + s[:, None] * hor + t[:, None] * ver - origin
    rays["time"] = rng.random(n).astype(np.float32)
    # secondary rays: from the oracle's hit points of (a resampling of) the camera rays
    idx = n_uni + rng.integers(0, n_cam, n_sec)
    prim, tt, normal, _, _, _ = oracle_scene.trace_closest(rays[idx], seed=1)
    ok = prim >= 0
    o = rays["origin"][idx].astype(np.float64) + np.where(ok, tt, 0.0)[:, None] * rays["direction"][idx].astype(np.float64)
    nrm = np.where(ok[:, None], normal, np.array([[0.0, 1.0, 0.0]]))
    nrm = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    rays["origin"][n_uni + n_cam:] = o
    rays["direction"][n_uni + n_cam:] = cosine_about(rng, nrm)
    return rays


def quantise(rgb_sum, spp):
    """Display for SampledColor (color.rs:56-65) in numpy."""
    x = np.where(np.isnan(rgb_sum), 0.0, rgb_sum) / spp
    g = np.clip(np.sqrt(np.maximum(x, 0.0)), 0.0, 0.999)
    return (256.0 * g).astype(np.int64)
