"""CPU: the plain-C oracle (oracle/vp_oracle.c) against the committed golden fixtures, which were produced by the
REFERENCE kernels (kernel/*.cl compiled in place, tests/golden/make_golden.py).  This is what pins the oracle."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import common
import oracle as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load(path):
    g = dict(np.load(path))
    p = O.Params()
    assert g["params"].size == C.sizeof(p)
    C.memmove(C.byref(p), g["params"].tobytes(), C.sizeof(p))
    return g, p


def test_fixtures_exist():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_fixture(port, path):
    g, p = load(path)
    raw = g["raw"]
    ch = port.raw2quad(raw, p.fmt, p.wq, p.hq, stale=7)
    for c in range(4):
        np.testing.assert_array_equal(ch[c], g[f"ch{c}"])
    flat = port.resampling(ch, p.fmt, p.wf, p.hf, p.model, p.max_robot_height, p.field_scale, p.off_x, p.off_y, p.sample_mode)
    np.testing.assert_array_equal(flat, g["flat"])
    grad = port.gradient_dot(flat, p.grad_offset)
    np.testing.assert_array_equal(grad, g["grad"])
    hor = port.sat_horizontal(grad)
    np.testing.assert_array_equal(hor, g["hor"])
    sat = port.sat_vertical(hor)
    np.testing.assert_array_equal(sat, g["sat"])
    circ = port.circle(sat, p.circle_radius)
    common.assert_float_images_equal(circ, g["circ"])
    m, counter = port.blob_list(flat, circ, p.circ_threshold, p.min_score, p.blob_radius, p.max_blobs)
    np.testing.assert_array_equal(counter, g["counter"])
    common.assert_matches_equal(m, g["matches"].view(O.MATCH_DTYPE), ordered=False)
    np.testing.assert_array_equal(port.quad2rgba(ch, p.fmt, p.sample_mode), g["quad_rgba"])
    np.testing.assert_array_equal(port.rgba2nv12(flat)[: g["nv12_flat"].size], g["nv12_flat"])
    np.testing.assert_array_equal(port.f2nv12(grad)[: g["nv12_grad"].size], g["nv12_grad"])
    np.testing.assert_array_equal(port.quad2nv12(ch, p.fmt, p.sample_mode)[: g["nv12_quad"].size], g["nv12_quad"])
    common.assert_float_images_equal(port.blob_score(flat, circ, p.circ_threshold, p.blob_radius), g["blob_score"])
    common.assert_float_images_equal(port.circularize(grad, 3, 5), g["circularize"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_whole_frame_entry_point(port, path):
    g, p = load(path)
    r = port.detect(g["raw"], p)
    np.testing.assert_array_equal(r["flat"], g["flat"])
    np.testing.assert_array_equal(r["sat"], g["sat"])
    np.testing.assert_array_equal(r["counter"], g["counter"])
    common.assert_matches_equal(r["matches"], g["matches"].view(O.MATCH_DTYPE), ordered=False)
    assert r["max_abs_sat"] == float(g["max_abs_sat"])
