"""Pins the oracle end to end against the reference's own published render (rest_of_your_life.png).

The PNG is the only result artefact the reference ships for this path (no tests, no vectors).  Its region
means (tests/golden/rest_of_your_life_regions.json, made by tests/golden/make_reference_regions.py) must be
reproduced by the oracle's Cornell render.  Tolerance: 6 % relative + 0.004 absolute per channel — the PNG
is a 100-spp render whose PPM->PNG conversion happened outside the repo.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _linear_image(rgb_sum, spp):
    """Per-pixel mean radiance as the 8-bit file stores it: NaN sum -> 0, gamma 2, clamp 0.999, 256 levels."""
    x = np.where(np.isnan(rgb_sum), 0.0, rgb_sum) / spp
    q = np.floor(256.0 * np.clip(np.sqrt(np.maximum(x, 0.0)), 0.0, 0.999))
    return (q / 256.0) ** 2


def test_cornell_region_means_match_reference_png(rt, oracle):
    with open(os.path.join(HERE, "golden", "rest_of_your_life_regions.json")) as f:
        golden = json.load(f)
    api = rt.api
    hs = api.HostScene("cornel_box", seed=1)
    osc = oracle.OracleScene(hs.desc)
    scale, spp = 2, 40  # 300x300: every golden region scaled by 1/2
    p = hs.params(width=600 // scale, spp=spp)
    img, _, st = osc.render(hs.camera(), p)
    assert st.paths == 300 * 300 * spp
    lin = _linear_image(img, spp)
    for name, g in golden["regions"].items():
        r0, r1 = (v // scale for v in g["rows"])
        c0, c1 = (v // scale for v in g["cols"])
        got = lin[r0:r1, c0:c1].mean(axis=(0, 1))
        want = np.array(g["mean_linear"])
        assert np.all(np.abs(got - want) <= 0.06 * want + 0.004), f"{name}: oracle {got} vs reference png {want}"
    # geometry pins: the 14-15 px black border and the black mirror face of the box
    rows = np.flatnonzero(lin.sum(axis=2).max(axis=1) > 0)
    first, last = golden["first_last_nonblack_row"]
    assert abs(rows[0] * scale - first) <= 2 and abs(rows[-1] * scale - last) <= 2
    g = golden["regions"]["box_front_face"]
    assert lin[g["rows"][0] // scale:g["rows"][1] // scale, g["cols"][0] // scale:g["cols"][1] // scale].max() == 0.0
    # rays per path: the closed box with a 15x emitter terminates after ~5 segments
    assert 4.5 < st.rays / st.paths < 5.6
