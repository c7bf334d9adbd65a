"""`rt1w_trace_closest` (through the C ABI) against the committed closest-hit vectors of tests/golden/closest_hit_vectors.npz
(the oracle's frozen answers on 1024 rays per scene; tests/golden/make_closest_hit_vectors.py): primitive ids bit-exact on
every ray not tagged as a tie, distance <= 1e-5 relative, normal <= 1e-5, (u, v) <= 2e-5, front_face equal - no oracle
code runs here.  Both tree layouts and both builders where the scene has a tree."""
import numpy as np
import pytest

from test_oracle_golden_vectors import KW, load_golden

pytestmark = pytest.mark.gpu


def _check(gsc, z, name, seed, label):
    rays = z[name + "/rays"]
    gp, gt, gn, gff, guv = gsc.trace_closest(rays, seed=seed)
    op, ot, on, off, ouv, amb = (z[name + "/" + k] for k in ("prim", "t", "normal", "front_face", "uv", "ambiguous"))
    keep = amb == 0
    bad = keep & (gp != op)
    assert not bad.any(), f"{label}: {bad.sum()} primitive-id mismatches, first at ray {np.flatnonzero(bad)[:5]}"
    hit = keep & (op >= 0)
    rel_t = np.abs(gt[hit].astype(np.float64) - ot[hit]) / np.maximum(np.abs(ot[hit]), 1e-30)
    assert rel_t.max() <= 1e-5, f"{label}: hit distance off by {rel_t.max():.3e} relative"
    assert np.abs(gn[hit].astype(np.float64) - on[hit]).max() <= 1e-5, label
    assert (gff[hit] == off[hit]).all(), label
    du = np.abs(guv[hit, 0].astype(np.float64) - ouv[hit, 0])
    du = np.minimum(du, 1.0 - du)  # u wraps at the atan2 branch cut (math.rs:69)
    dv = np.abs(guv[hit, 1].astype(np.float64) - ouv[hit, 1])
    assert max(du.max(), dv.max()) <= 2e-5, label
    assert np.isinf(gt[keep & (op < 0)]).all(), label


@pytest.mark.parametrize("name", load_golden()[1])
def test_closest_hits_match_golden_vectors(rt, gpu_ctx, monkeypatch, name):
    api = rt.api
    z, _, seed = load_golden()
    hs = api.HostScene(name, seed=1, **KW.get(name, {}))
    gsc = api.Scene(gpu_ctx, hs.desc)
    _check(gsc, z, name, seed, name)
    gsc.close()
    if name not in ("random_scene", "final_scene", "stress"):  # the small scenes are scanned, not walked
        return
    for env, value in (("RT1W_BVH_LAYOUT", "wide"), ("RT1W_BVH_LAYOUT", "binary"), ("RT1W_BVH_BUILDER", "lbvh"), ("RT1W_BVH_BUILDER", "sah")):
        monkeypatch.setenv(env, value)
        gsc = api.Scene(gpu_ctx, hs.desc)
        monkeypatch.delenv(env)
        _check(gsc, z, name, seed, f"{name}, {env}={value}")
        gsc.close()
