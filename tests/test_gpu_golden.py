"""GPU: libvp_b200.so against the committed golden fixtures produced by the reference kernels."""
import numpy as np
import pytest

import common
import oracle as O
from test_golden import GOLDEN, load
import os

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_matches_reference_fixture(ctx, path):
    g, p = load(path)
    got = ctx.detect(g["raw"], common.to_vp(p))
    np.testing.assert_array_equal(got["flat"], g["flat"])
    np.testing.assert_array_equal(got["grad"], g["grad"])
    common.assert_float_images_equal(got["circ"], g["circ"])
    np.testing.assert_array_equal(got["counter"][0], g["counter"])
    common.assert_matches_equal(got["matches"][0], g["matches"].view(O.MATCH_DTYPE), ordered=False)
    # stage API on the fixture's intermediates
    ch = [g[f"ch{c}"] for c in range(4)]
    got_ch = ctx.raw2quad(g["raw"], p.fmt, p.wq, p.hq, stale=7)
    for c in range(4):
        np.testing.assert_array_equal(got_ch[c], ch[c])
    np.testing.assert_array_equal(ctx.sat_horizontal(g["grad"]), g["hor"])
    np.testing.assert_array_equal(ctx.sat_vertical(g["hor"]), g["sat"])
    np.testing.assert_array_equal(ctx.quad2rgba(ch, p.fmt, p.sample_mode), g["quad_rgba"])
    np.testing.assert_array_equal(ctx.rgba2nv12(g["flat"])[: g["nv12_flat"].size], g["nv12_flat"])
    np.testing.assert_array_equal(ctx.f2nv12(g["grad"])[: g["nv12_grad"].size], g["nv12_grad"])
    if p.wq % 2 == 0 and p.hq % 2 == 0:
        np.testing.assert_array_equal(ctx.quad2nv12(ch, p.fmt, p.sample_mode)[: g["nv12_quad"].size], g["nv12_quad"])
    common.assert_float_images_equal(ctx.blob_score(g["flat"], g["circ"], p.circ_threshold, p.blob_radius), g["blob_score"])
    common.assert_float_images_equal(ctx.circularize(g["grad"], 3, 5), g["circularize"])
