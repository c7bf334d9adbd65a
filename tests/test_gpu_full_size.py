"""BASELINE config 1 at FULL size (600x600, 100 spp, depth 50) on the GPU, checked through what does not need a
CPU render of that size: the reference's own published image, and size-independent properties of the estimator."""
import json
import os

import numpy as np
import pytest

from test_oracle_golden_image import _linear_image

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def full(rt, gpu_ctx):
    api = rt.api
    hs = api.HostScene("cornel_box", seed=1)
    gsc = api.Scene(gpu_ctx, hs.desc)
    p = hs.params()  # the arm's own settings: 600 x 600, 100 spp, depth 50 (main.rs:867-894)
    assert (p.width, p.height, p.sample_end, p.max_depth) == (600, 600, 100, 50)
    img, _, st = gsc.render(hs.camera(), p)
    yield api, hs, gsc, img, st
    gsc.close()


def test_region_means_match_reference_png(full):
    """The product against the reference's only published result (rest_of_your_life.png), same bar as the oracle's."""
    api, hs, gsc, img, st = full
    with open(os.path.join(HERE, "golden", "rest_of_your_life_regions.json")) as f:
        golden = json.load(f)
    assert st.paths == 600 * 600 * 100 and 4.5 < st.rays / st.paths < 5.6
    lin = _linear_image(img, 100)
    for name, g in golden["regions"].items():
        got = lin[g["rows"][0]:g["rows"][1], g["cols"][0]:g["cols"][1]].mean(axis=(0, 1))
        want = np.array(g["mean_linear"])
        assert np.all(np.abs(got - want) <= 0.06 * want + 0.004), f"{name}: gpu {got} vs reference png {want}"
    rows = np.flatnonzero(lin.sum(axis=2).max(axis=1) > 0)
    first, last = golden["first_last_nonblack_row"]
    assert abs(rows[0] - first) <= 2 and abs(rows[-1] - last) <= 2
    g = golden["regions"]["box_front_face"]
    assert lin[g["rows"][0]:g["rows"][1], g["cols"][0]:g["cols"][1]].max() == 0.0  # the mirror face: NaN sums resolve to black


def test_eight_sample_ranges_add_up_at_full_size(full):
    """The multi-GPU split of config 4 (sample ranges, one per rank) at config 1's size: 8 shards == 1 render."""
    api, hs, gsc, img, st = full
    cam = hs.camera()
    total = np.zeros_like(img)
    rays = 0
    for r in range(8):
        s0, s1 = (100 * r) // 8, (100 * (r + 1)) // 8
        part, _, ps = gsc.render(cam, hs.params(spp=s1, sample_begin=s0))
        total += part
        rays += ps.rays
    assert rays == st.rays
    ok = np.isfinite(img) & np.isfinite(total)
    assert (np.isfinite(img) == np.isfinite(total)).all()
    assert np.allclose(total[ok], img[ok], rtol=2e-4, atol=2e-4)


def test_device_image_is_the_resolved_render(full):
    """rt1w_render_rgb8 at full size against the host resolve of the fp32 sums (summation order may move a level)."""
    api, hs, gsc, img, st = full
    rgb8, _ = gsc.render_rgb8(hs.camera(), hs.params())
    want = api.resolve_rgb8(img, 100).astype(np.int64)
    assert (np.abs(rgb8.astype(np.int64) - want) <= 1).mean() >= 0.9999
