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


def make_ray_set(api, host_scene, oracle_scene, prims, n, max_extent=None, aspect=None):
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
    cam = host_scene.camera() if aspect is None else host_scene.camera(aspect=aspect)  # aspect: what Camera::new is given (camera.rs:27)
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


def check_trace_parity(gsc, osc, rays, seed=0x5EED, max_ambiguous=0.02, min_hit_fraction=0.1, label="", cache=None):
    """north_star checks 1 and 2 on one ray set: primitive ids bit-exact on every ray the oracle does not tag as a tie,
    hit distance <= 1e-5 relative, normal <= 1e-5, (u, v) <= 2e-5, front_face equal, misses report t = inf.
    Prints the ambiguous fraction (pytest -s / -rP shows it) and returns it.
    cache: a dict that keeps the oracle's answers between calls on the same rays (several device scenes, one oracle)."""
    n = len(rays)
    gp, gt, gn, gff, guv = gsc.trace_closest(rays, seed=seed)
    if cache is not None and "trace" in cache:
        op, ot, on, off, ouv, amb = cache["trace"]
    else:
        op, ot, on, off, ouv, amb = osc.trace_closest(rays, seed=seed)
        if cache is not None:
            cache["trace"] = (op, ot, on, off, ouv, amb)
    keep = amb == 0
    frac_amb = 1.0 - keep.mean()
    print(f"[trace parity] {label}: {n} rays, {frac_amb:.5f} ambiguous (ties within 1e-9 relative), {(op >= 0).mean():.3f} hit")
    assert frac_amb <= max_ambiguous, f"{frac_amb:.4f} of the rays are ambiguous: the ray set is badly conditioned"
    bad = keep & (gp != op)
    assert not bad.any(), f"{bad.sum()} primitive-id mismatches, first at ray {np.flatnonzero(bad)[:5]}: gpu {gp[bad][:5]} oracle {op[bad][:5]}"
    hit = keep & (op >= 0)
    assert hit.sum() >= min_hit_fraction * n
    rel_t = np.abs(gt[hit].astype(np.float64) - ot[hit]) / np.maximum(np.abs(ot[hit]), 1e-30)
    assert rel_t.max() <= 1e-5, f"hit distance off by {rel_t.max():.3e} relative"
    dn = np.abs(gn[hit].astype(np.float64) - on[hit]).max()
    assert dn <= 1e-5, f"normal off by {dn:.3e}"
    assert (gff[hit] == off[hit]).all()
    du = np.abs(guv[hit, 0].astype(np.float64) - ouv[hit, 0])
    du = np.minimum(du, 1.0 - du)  # u wraps at the atan2 branch cut (math.rs:69)
    dv = np.abs(guv[hit, 1].astype(np.float64) - ouv[hit, 1])
    assert max(du.max(), dv.max()) <= 2e-5
    miss = keep & (op < 0)
    assert np.isinf(gt[miss]).all()
    return frac_amb


def _stats(stat, n):
    mean = stat[..., :3] / n
    var = np.maximum(stat[..., 3:] / n - mean * mean, 0.0)
    return mean, var


def check_render_parity(api, gsc, osc, cam, make_params, spp, stat_clamp=20.0, flags=0, cache=None):
    """north_star check 3 (different RNG streams, so statistical): >= 99.9 % of the pixel-channel means within 5 sigma of
    the oracle's, RMSE(gpu, oracle) <= 1.15 RMSE(oracle, oracle') + 0.5 8-bit level, image mean and rays/path within 2 %.
    make_params(spp, sample_begin, flags, stat_clamp, seed) -> RenderParams."""
    p = make_params(spp, 0, api.FLAG_STATS | flags, stat_clamp, 3)
    g_sum, g_stat, g_st = gsc.render(cam, p, want_stat=True)
    assert g_st.paths == p.width * p.height * spp
    if cache is not None and "render" in cache:  # several device renders (kernels, builders) against one pair of oracle images
        oa_stat, oa_st, ob_stat = cache["render"]
    else:
        oa_sum, oa_stat, oa_st = osc.render(cam, p, want_stat=True)
        pb = make_params(2 * spp, spp, api.FLAG_STATS, stat_clamp, 0)
        ob_sum, ob_stat, _ = osc.render(cam, pb, want_stat=True)
        if cache is not None:
            cache["render"] = (oa_stat, oa_st, ob_stat)
    gm, gv = _stats(g_stat.astype(np.float64), spp)
    am, av = _stats(oa_stat, spp)
    bm, _ = _stats(ob_stat, spp)
    sigma = np.sqrt((gv + av) / spp) + 1e-4
    z = np.abs(gm - am) / sigma
    frac_ok = (z <= 5.0).mean()
    assert frac_ok >= 0.999, f"only {frac_ok:.5f} of the pixel means are within 5 sigma (worst z = {z.max():.1f})"
    img_g, img_a, img_b = (quantise(m, 1) for m in (gm, am, bm))
    rmse_ga = np.sqrt(np.mean((img_g - img_a) ** 2.0))
    rmse_ab = np.sqrt(np.mean((img_a - img_b) ** 2.0))
    assert rmse_ga <= 1.15 * rmse_ab + 0.5, f"RMSE gpu-oracle {rmse_ga:.2f} vs oracle-oracle {rmse_ab:.2f} (8-bit levels)"
    # whole-image mean within 2 % (+ 4 sigma of the two estimates: scenes lit by rarely-found lights, like the stress scene,
    # have image means that are themselves noisy at test sizes)
    sigma_mean = np.sqrt((gv + av).sum() / spp) / gv.size
    assert abs(gm.mean() - am.mean()) <= 0.02 * am.mean() + 4.0 * sigma_mean + 1e-4, \
        f"image mean gpu {gm.mean():.5f} oracle {am.mean():.5f} (sigma of the difference {sigma_mean:.5f})"
    rpp_g, rpp_o = g_st.rays / g_st.paths, oa_st.rays / oa_st.paths
    assert abs(rpp_g - rpp_o) <= 0.02 * rpp_o, f"rays/path gpu {rpp_g:.3f} oracle {rpp_o:.3f}"
    return g_st, oa_st


def check_same_render(a, b, label=""):
    """Two renders of the same paths by two differently compiled kernels (lockstep / persistent, binary / 8-wide): the same
    closest hits, so the same image - except that the two instantiations may contract an f64 expression differently, and a
    last-bit difference in one hit distance sends a deep path (depth up to 50) another way: a handful of rays per million.
    a, b: (rgb_sum, stat, RenderStats)."""
    (ia, _, sa), (ib, _, sb) = a, b
    assert sa.paths == sb.paths, label
    assert abs(float(sa.rays) - float(sb.rays)) <= 1e-4 * float(sa.rays), f"{label}: {sa.rays} vs {sb.rays} rays"
    ok = np.isfinite(ia) & np.isfinite(ib)
    close = np.isclose(ia, ib, rtol=1e-3, atol=1e-3) | ~ok
    assert close.all(axis=2).mean() >= 0.999, f"{label}: {(~close.all(axis=2)).sum()} pixels differ"
    assert abs(np.nansum(ia) - np.nansum(ib)) <= 2e-3 * abs(np.nansum(ia)), label
