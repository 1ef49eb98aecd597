"""CPU: the plain-C restatement against the reference's own kernel sources (compiled in place through oracle/clemu.h)
on random and adversarial inputs.  Skipped where oracle/_ref was not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

import common


def rand_rgba(rng, h, w):
    return rng.integers(0, 256, (h, w, 4), dtype=np.uint8)


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_raw2quad_resampling_quad_conversions(port, clref, fmt, mode):
    p, raw, _ = common.make_case(wq=70, hq=46, fmt=fmt, k2=0.09, tilt=0.17, seed=fmt * 3 + mode, sample_mode=mode)
    a, b = port.raw2quad(raw, fmt, p.wq, p.hq, stale=9), clref.raw2quad(raw, fmt, p.wq, p.hq, stale=9)
    for c in range(4):
        np.testing.assert_array_equal(a[c], b[c])
    args = (a, fmt, p.wf, p.hf, p.model, p.max_robot_height, p.field_scale, p.off_x, p.off_y, mode)
    np.testing.assert_array_equal(port.resampling(*args), clref.resampling(*args))
    # far outside the sensor: every tap clamps
    args2 = (a, fmt, 40, 30, p.model, 180.0, 11.0, -3000.0, 2500.0, mode)
    np.testing.assert_array_equal(port.resampling(*args2), clref.resampling(*args2))
    np.testing.assert_array_equal(port.quad2rgba(a, fmt, mode), clref.quad2rgba(a, fmt, mode))
    n = p.wq * p.hq * 3 // 2
    np.testing.assert_array_equal(port.quad2nv12(a, fmt, mode)[:n], clref.quad2nv12(a, fmt, mode)[:n])


@pytest.mark.parametrize("offset", [0, 1, 3])
def test_gradient_sat_circle(port, clref, offset):
    rng = np.random.default_rng(offset)
    img = rand_rgba(rng, 41, 57)
    g = port.gradient_dot(img, offset)
    np.testing.assert_array_equal(g, clref.gradient_dot(img, offset))
    f = (rng.standard_normal((41, 57)) * 1e5).astype(np.float32)  # arbitrary floats: sequential rounding must agree
    for x in (g, f):
        h = port.sat_horizontal(x)
        np.testing.assert_array_equal(h, clref.sat_horizontal(x))
        v = port.sat_vertical(h)
        np.testing.assert_array_equal(v, clref.sat_vertical(h))
        for r in (0, 1, 4, 7):
            common.assert_float_images_equal(port.circle(v, r), clref.circle(v, r))


@pytest.mark.parametrize("min_score", [0.0, 0.7, -2.0])
def test_blob_list_and_dead_kernels(port, clref, min_score):
    rng = np.random.default_rng(17)
    h, w = 48, 66
    img = rand_rgba(rng, h, w)
    circ = (rng.standard_normal((h, w)) * 20).astype(np.float32)
    circ[5:8, 9:14] = 77.0
    for radius, mx in [(0, 500), (3, 500), (5, 9)]:
        a, ca = port.blob_list(img, circ, 15.0, min_score, radius, mx)
        b, cb = clref.blob_list(img, circ, 15.0, min_score, radius, mx)
        np.testing.assert_array_equal(ca, cb)
        common.assert_matches_equal(a, b, ordered=True)  # both run work-items in raster order
    common.assert_float_images_equal(port.blob_score(img, circ, 15.0, 4), clref.blob_score(img, circ, 15.0, 4))
    g = rng.integers(-9000, 9000, (h, w)).astype(np.float32)
    common.assert_float_images_equal(port.circularize(g, 2, 6), clref.circularize(g, 2, 6))


def test_nv12(port, clref):
    rng = np.random.default_rng(4)
    img = rand_rgba(rng, 20, 34)
    n = 20 * 34 * 3 // 2
    np.testing.assert_array_equal(port.rgba2nv12(img)[:n], clref.rgba2nv12(img)[:n])
    f = (rng.standard_normal((20, 34)) * 300).astype(np.float32)
    f[0, :3] = [np.nan, np.inf, -np.inf]
    np.testing.assert_array_equal(port.f2nv12(f)[:n], clref.f2nv12(f)[:n])


def test_whole_frame(port, clref):
    for kw in [dict(wq=96, hq=64), dict(wq=120, hq=90, fmt=1, k2=0.1, tilt=0.25, n_robots=2), dict(wq=64, hq=48, frame="noise", max_blobs=30)]:
        p, raw, _ = common.make_case(**kw)
        a, b = port.detect(raw, p), clref.detect(raw, p)
        for k in ("flat", "grad", "sat"):
            np.testing.assert_array_equal(a[k], b[k])
        common.assert_float_images_equal(a["circ"], b["circ"])
        np.testing.assert_array_equal(a["counter"], b["counter"])
        common.assert_matches_equal(a["matches"], b["matches"])
        assert a["max_abs_sat"] == b["max_abs_sat"]
