"""Pins the oracle end to end against the reference's own published render (rest_of_your_life.png).

The PNG is the only result artefact the reference ships for this path (no tests, no vectors): ONE 100-spp realisation of
main.rs scene 5 at 600x600.  The oracle renders the same configuration (another realisation: different RNG glue, see
DESIGN.md section 2) and both go through the file's own quantisation (color.rs:14-21,56-65, inverted as (v / 256)^2).
tests/golden/rest_of_your_life_regions.json holds the PNG's region means (tests/golden/make_reference_regions.py) and,
per region and channel, the standard deviation sigma of such a mean over eight oracle realisations
(tests/golden/make_oracle_sigma.py).  Two independent realisations differ by sqrt(2) sigma; the bar is 4 sqrt(2) sigma -
about 0.5 % of a region mean, where round 1 allowed 6 %.  Further pins: the black frame (rows and columns), the share of
pure-black pixels inside it (the mirror face of the box: NaN sums resolve to black), the saturated light.
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


def load_golden():
    with open(os.path.join(HERE, "golden", "rest_of_your_life_regions.json")) as f:
        return json.load(f)


def check_against_reference_png(lin, golden, who):
    """`lin`: a 600x600 render of 100 spp through _linear_image.  The bars both the oracle and the GPU render must meet."""
    assert lin.shape == (600, 600, 3)
    worst = 0.0
    for name, g in golden["regions"].items():
        got = lin[g["rows"][0]:g["rows"][1], g["cols"][0]:g["cols"][1]].mean(axis=(0, 1))
        want, sigma = np.array(g["mean_linear"]), np.array(g["oracle_sigma_linear"])
        bar = 4.0 * np.sqrt(2.0) * sigma + 2e-5  # (2e-5: the JSON keeps six decimals)
        assert np.all(np.abs(got - want) <= bar), f"{name}: {who} {got} vs reference png {want}, allowed {bar}"
        worst = max(worst, float(np.max(np.abs(got - want) / np.maximum(want, 1e-9))) if want.max() > 0 else 0.0)
    print(f"[reference png] {who}: every region mean within 4 sqrt(2) sigma; largest relative difference {worst:.4%}")
    # the black frame: rows 0..14 and 587..599, columns 0..13 and 586..599 of the published image are black
    rows = np.flatnonzero(lin.sum(axis=2).max(axis=1) > 0)
    cols = np.flatnonzero(lin.sum(axis=2).max(axis=0) > 0)
    assert [int(rows[0]), int(rows[-1])] == golden["first_last_nonblack_row"], (rows[0], rows[-1])
    assert [int(cols[0]), int(cols[-1])] == golden["first_last_nonblack_col"], (cols[0], cols[-1])
    # pure-black pixels inside the frame (6.5 %: the mirror face, whose NaN sums resolve to black, color.rs:16-18)
    (r0, r1), (c0, c1) = golden["first_last_nonblack_row"], golden["first_last_nonblack_col"]
    black = float((lin[r0:r1 + 1, c0:c1 + 1].max(axis=2) == 0.0).mean())
    assert abs(black - golden["black_fraction_inside_frame"]) <= 4.0 * np.sqrt(2.0) * golden["oracle_black_fraction"][1] + 1e-4, black
    g = golden["regions"]["box_front_face"]
    assert lin[g["rows"][0]:g["rows"][1], g["cols"][0]:g["cols"][1]].max() == 0.0
    g = golden["regions"]["light"]  # every value on the light is the top level, 255
    assert (lin[g["rows"][0]:g["rows"][1], g["cols"][0]:g["cols"][1]] >= (255.0 / 256.0) ** 2).all()
    assert golden["light_saturated_fraction"] == 1.0


def test_cornell_render_matches_reference_png(rt, oracle):
    golden = load_golden()
    api = rt.api
    hs = api.HostScene("cornel_box", seed=1)
    osc = oracle.OracleScene(hs.desc)
    p = hs.params()  # the arm's own settings (main.rs:867-894): 600 x 600, 100 spp, depth 50
    assert (p.width, p.height, p.sample_end, p.max_depth) == (600, 600, 100, 50)
    p.sample_begin, p.sample_end = 800, 900  # a realisation the sigma estimate has not seen
    img, _, st = osc.render(hs.camera(), p)
    assert st.paths == 600 * 600 * 100
    check_against_reference_png(_linear_image(img, 100), golden, "oracle")
    # rays per path: the closed box with a 15x emitter terminates after ~5 segments
    assert 4.9 < st.rays / st.paths < 5.3
