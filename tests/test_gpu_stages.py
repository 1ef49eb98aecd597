"""GPU parity, stage by stage: every reference kernel's CUDA replacement against the CPU oracle, bit-exact.

All calls go through the C ABI (ctypes -> libvp_b200.so).  The oracle (oracle/) is only the checker."""
import numpy as np
import pytest

import common
import oracle as O
from vpb200 import lib

pytestmark = pytest.mark.gpu

CASES = [
    dict(wq=96, hq=64, fmt=0),
    dict(wq=102, hq=66, fmt=0, k2=0.12, tilt=0.2),      # width not a multiple of 4: scalar paths
    dict(wq=96, hq=64, fmt=1, k2=-0.08, tilt=-0.1),
    dict(wq=128, hq=96, fmt=2, k2=0.05),
    dict(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.3, n_robots=4, n_balls=3, seed=7),
]


def planes(port, p, raw, stale=0):
    return port.raw2quad(raw, p.fmt, p.wq, p.hq, stale=stale)


@pytest.mark.parametrize("kw", CASES)
def test_raw2quad(ctx, port, kw):
    p, raw, _ = common.make_case(**kw)
    want = port.raw2quad(raw, p.fmt, p.wq, p.hq, stale=77)
    got = ctx.raw2quad(raw, p.fmt, p.wq, p.hq, stale=77)
    for c in range(4):  # BGR leaves plane 3 untouched (raw2quad.cl:23-29)
        np.testing.assert_array_equal(got[c], want[c])


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("kw", CASES)
def test_resampling(ctx, port, kw, mode):
    p, raw, lp = common.make_case(**kw)
    ch = planes(port, p, raw)
    want = port.resampling(ch, p.fmt, p.wf, p.hf, p.model, p.max_robot_height, p.field_scale, p.off_x, p.off_y, mode)
    got = ctx.resampling(ch, p.fmt, p.wf, p.hf, lib.camera_model_from_bytes(lp.model_bytes), p.max_robot_height, p.field_scale,
                         p.off_x, p.off_y, mode)
    np.testing.assert_array_equal(got, want)


def test_resampling_out_of_image(ctx, port):
    """Flat pixels that project outside the sensor: CLAMP_TO_EDGE on every tap."""
    p, raw, lp = common.make_case(wq=64, hq=48)
    ch = planes(port, p, raw)
    m = lib.camera_model_from_bytes(lp.model_bytes)
    for (ox, oy, s) in [(-2000.0, -1500.0, 9.0), (500.0, 400.0, 3.0), (-100.0, 300.0, 0.5)]:
        want = port.resampling(ch, p.fmt, 80, 60, p.model, 180.0, s, ox, oy, 0)
        got = ctx.resampling(ch, p.fmt, 80, 60, m, 180.0, s, ox, oy, 0)
        np.testing.assert_array_equal(got, want)


def rand_rgba(rng, h, w):
    a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    a[..., 3] = 255
    return a


@pytest.mark.parametrize("shape", [(64, 96), (37, 131), (2, 2), (1, 5), (200, 1224)])
@pytest.mark.parametrize("offset", [0, 1, 2, 5])
def test_gradient_dot(ctx, port, shape, offset):
    rng = np.random.default_rng(offset * 100 + shape[0])
    a = rand_rgba(rng, *shape)
    np.testing.assert_array_equal(ctx.gradient_dot(a, offset), port.gradient_dot(a, offset))


def test_gradient_dot_ignores_alpha(ctx, port):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (40, 52, 4), dtype=np.uint8)  # arbitrary alpha: gradientDot.cl:29 sums x, y, z only
    np.testing.assert_array_equal(ctx.gradient_dot(a, 2), port.gradient_dot(a, 2))


@pytest.mark.parametrize("shape", [(64, 96), (37, 131), (1, 1), (3, 700), (600, 5)])
def test_sat_sequential_any_float(ctx, port, shape):
    """The stage API scans in the reference's order, so it is bit-exact for arbitrary fp32 input."""
    rng = np.random.default_rng(shape[1])
    a = (rng.standard_normal(shape) * 1e4).astype(np.float32)
    hor_w = port.sat_horizontal(a)
    np.testing.assert_array_equal(ctx.sat_horizontal(a), hor_w)
    np.testing.assert_array_equal(ctx.sat_vertical(hor_w), port.sat_vertical(hor_w))


@pytest.mark.parametrize("r", [0, 1, 2, 5, 9])
def test_circle(ctx, port, r):
    rng = np.random.default_rng(r)
    g = rng.integers(-195075, 195076, (70, 90)).astype(np.float32)
    sat = port.sat_vertical(port.sat_horizontal(g))
    got, want = ctx.circle(sat, r), port.circle(sat, r)
    common.assert_float_images_equal(got, want)


def circ_field(rng, h, w, plateau=False):
    c = (rng.standard_normal((h, w)) * 20).astype(np.float32)
    if plateau:
        c[10:14, 20:26] = 99.0   # plateau: every pixel of it is a "peak" with NaN offsets (blobList.cl:93-94)
        c[0, 0] = 200.0          # border peaks: clamped neighbours equal the centre
        c[h - 1, w - 1] = 150.0
    return c


@pytest.mark.parametrize("radius", [0, 1, 4, 6])
@pytest.mark.parametrize("min_score", [0.0, 0.8, -1.0])
def test_blob_list(ctx, port, radius, min_score):
    rng = np.random.default_rng(radius)
    h, w = 60, 83
    rgba, circ = rand_rgba(rng, h, w), circ_field(rng, h, w, plateau=True)
    want, wc = port.blob_list(rgba, circ, 15.0, min_score, radius, 2000)
    got, gc = ctx.blob_list(rgba, circ, 15.0, min_score, radius, 2000)
    np.testing.assert_array_equal(gc, wc)
    common.assert_matches_equal(got, want, ordered=True)  # raster order, like the sequential oracle
    assert len(want) >= (10 if min_score <= 0 else 1)


def test_blob_list_overflow_and_negative_threshold(ctx, port):
    rng = np.random.default_rng(3)
    h, w = 50, 70
    rgba, circ = rand_rgba(rng, h, w), circ_field(rng, h, w)
    for thr, mx in [(-1e9, 25), (5.0, 7), (5.0, 0), (float("nan"), 40)]:
        want, wc = port.blob_list(rgba, circ, thr, 0.0, 3, mx)
        got, gc = ctx.blob_list(rgba, circ, thr, 0.0, 3, mx)
        np.testing.assert_array_equal(gc, wc)            # counter[0] keeps counting past max (blobList.cl:87-89)
        common.assert_matches_equal(got, want)
        assert wc[0] > mx


def test_blob_list_appends_after_existing_counter(ctx, port):
    """The kernel increments the counters it is given (main.cpp zeroes them; a caller may not)."""
    rng = np.random.default_rng(11)
    rgba, circ = rand_rgba(rng, 40, 40), circ_field(rng, 40, 40)
    want, wc = port.blob_list(rgba, circ, 15.0, 0.0, 3, 500)
    got, gc = ctx.blob_list(rgba, circ, 15.0, 0.0, 3, 500, counter0=(5, 2, 9))
    np.testing.assert_array_equal(gc, wc + np.array([5, 2, 9]))
    common.assert_matches_equal(got, want[: len(got)])
    assert len(got) == len(want)


@pytest.mark.parametrize("shape", [(64, 96), (2, 2), (30, 130)])
def test_rgba2nv12_and_f2nv12(ctx, port, shape):
    rng = np.random.default_rng(shape[1])
    a = rand_rgba(rng, *shape)
    n = shape[0] * shape[1] * 3 // 2
    np.testing.assert_array_equal(ctx.rgba2nv12(a)[:n], port.rgba2nv12(a)[:n])
    f = (rng.standard_normal(shape) * 200).astype(np.float32)
    f[0, 0], f[0, 1], f[1, 0] = np.nan, np.inf, -np.inf
    np.testing.assert_array_equal(ctx.f2nv12(f)[:n], port.f2nv12(f)[:n])


def test_nv12_rejects_odd_sizes(ctx):
    with pytest.raises(lib.VpError):
        ctx.rgba2nv12(np.zeros((5, 6, 4), np.uint8))


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("kw", CASES[:4])
def test_quad2nv12_quad2rgba(ctx, port, kw, mode):
    p, raw, _ = common.make_case(**kw)
    ch = planes(port, p, raw)
    np.testing.assert_array_equal(ctx.quad2rgba(ch, p.fmt, mode), port.quad2rgba(ch, p.fmt, mode))
    if p.wq % 2 == 0 and p.hq % 2 == 0:
        n = p.wq * p.hq * 3 // 2
        np.testing.assert_array_equal(ctx.quad2nv12(ch, p.fmt, mode)[:n], port.quad2nv12(ch, p.fmt, mode)[:n])
        # the fused variants read the raw frame directly
        np.testing.assert_array_equal(ctx.raw2nv12(raw, p.fmt, p.wq, p.hq, mode)[:n], port.quad2nv12(ch, p.fmt, mode)[:n])
    np.testing.assert_array_equal(ctx.raw2rgba(raw, p.fmt, p.wq, p.hq, mode), port.quad2rgba(ch, p.fmt, mode))


def test_dead_kernels(ctx, port):
    rng = np.random.default_rng(2)
    g = rng.integers(-5000, 5000, (40, 56)).astype(np.float32)
    got, want = ctx.circularize(g, 3, 5), port.circularize(g, 3, 5)
    common.assert_float_images_equal(got, want)
    rgba, circ = rand_rgba(rng, 40, 56), circ_field(rng, 40, 56, plateau=True)
    got, want = ctx.blob_score(rgba, circ, 15.0, 4), port.blob_score(rgba, circ, 15.0, 4)
    common.assert_float_images_equal(got, want)


def test_errors_are_reported_not_fatal(ctx):
    with pytest.raises(lib.VpError) as e:
        ctx.gradient_dot(np.zeros((4, 4), np.uint8), 1)  # U8 image where RGBA8 is required
    assert e.value.code == 1
    assert "RGBA8" in str(e.value)


def test_batched_nv12_views_match_the_per_frame_calls(ctx, port):
    """The three debug-stream conversions over a batch in one launch (main.cpp:380-393's views for n frames), with a frame
    stride larger than the 1.5*w*h that is written."""
    rng = np.random.default_rng(5)
    n, w, h = 5, 46, 30
    rgba = rng.integers(0, 256, (n, h, w, 4), dtype=np.uint8)
    f32 = (rng.standard_normal((n, h, w)) * 120).astype(np.float32)
    used = w * h * 3 // 2
    got = ctx.nv12_batch("rgba", rgba, w, h, stride=used + 64)
    for i in range(n):
        np.testing.assert_array_equal(got[i, :used], port.rgba2nv12(rgba[i])[:used])
        assert not got[i, used:].any()                                   # nothing is written past 1.5*w*h
    got = ctx.nv12_batch("f32", f32, w, h)
    for i in range(n):
        np.testing.assert_array_equal(got[i, :used], port.f2nv12(f32[i])[:used])
    for kw in (dict(wq=48, hq=32, fmt=0), dict(wq=46, hq=30, fmt=1, k2=0.1), dict(wq=40, hq=24, fmt=2)):
        frames = []
        for s_ in range(3):
            p, raw, _ = common.make_case(seed=40 + s_, **kw)
            frames.append(raw)
        used = p.wq * p.hq * 3 // 2
        got = ctx.nv12_batch("raw", np.stack(frames), p.wq, p.hq, fmt=p.fmt)
        for i, raw in enumerate(frames):
            ch = port.raw2quad(raw, p.fmt, p.wq, p.hq)
            np.testing.assert_array_equal(got[i, :used], port.quad2nv12(ch, p.fmt, 0)[:used])
    with pytest.raises(lib.VpError):
        ctx.nv12_batch("rgba", rgba, w, h, stride=used - 2)              # stride smaller than a frame


def test_batched_nv12_surfaces_are_what_an_encoder_reads(ctx, port):
    """the planes vp_nv12_surface_of describes for frame i of a batch ARE frame i's luma and chroma --
    read the way the reference's encoder thread reads them (rtpstreamer.cpp:120-121,177-181: rows of `pitch` bytes from `y`,
    height/2 rows of interleaved U,V from `uv`) -- and they never leave the device."""
    import ctypes as C
    rng = np.random.default_rng(11)
    n, w, h = 4, 64, 48
    rgba = rng.integers(0, 256, (n, h, w, 4), dtype=np.uint8)
    stride = 2 * w * h
    src = ctx.buffer(rgba.nbytes, rgba.reshape(-1))
    dst = ctx.buffer(n * stride)
    ctx._ck(ctx.lib.vp_rgba2nv12_batch_device(ctx.h, C.c_void_p(src.device_ptr), n, w, h, C.c_void_p(dst.device_ptr), stride))
    for i in range(n):
        s = lib.Nv12Surface()
        ctx._ck(ctx.lib.vp_nv12_surface_of(C.c_void_p(dst.device_ptr), w, h, stride, i, C.byref(s)))
        assert s.aligned16 == 1 and s.pitch_y == w and s.pitch_uv == w
        y = ctx.to_host(s.y, s.pitch_y * h).reshape(h, s.pitch_y)[:, :w]
        uv = ctx.to_host(s.uv, s.pitch_uv * h // 2).reshape(h // 2, s.pitch_uv)[:, :w]
        want = port.rgba2nv12(rgba[i])
        np.testing.assert_array_equal(y.reshape(-1), want[: w * h])
        np.testing.assert_array_equal(uv.reshape(-1), want[w * h: w * h * 3 // 2])
    src.release()
    dst.release()


@pytest.mark.parametrize("fmt", [0, 1])
def test_wide_nv12_kernels_on_random_bytes(ctx, port, fmt):
    """The 8x2-pixels-per-thread kernels (w % 8 == 0, aligned views): uniformly random bytes hit every rounding case of the
    integer 16-bit-lane demosaic (ties of the 1/16 grid, both parities), the clamped left column and top row, and a
    frame stride that leaves a gap."""
    rng = np.random.default_rng(17 + fmt)
    n, wq, hq = 3, 72, 46
    raws = rng.integers(0, 256, (n, 4 * wq * hq), dtype=np.uint8)
    raws[0, : 8 * wq] = np.tile(np.array([0, 255, 255, 0], np.uint8), 2 * wq)   # extremes next to each other
    used = wq * hq * 3 // 2
    stride = used + 72
    got = ctx.nv12_batch("raw", raws, wq, hq, fmt=fmt, stride=stride)
    for i in range(n):
        ch = port.raw2quad(raws[i], fmt, wq, hq)
        np.testing.assert_array_equal(got[i, :used], port.quad2nv12(ch, fmt, 0)[:used])
        assert not got[i, used:].any()
        np.testing.assert_array_equal(ctx.raw2nv12(raws[i], fmt, wq, hq, 0)[:used], got[i, :used])
        np.testing.assert_array_equal(ctx.raw2rgba(raws[i], fmt, wq, hq, 0), port.quad2rgba(ch, fmt, 0))
        odd = raws[i][: 4 * wq * (hq - 1)]                                       # an odd number of quad rows
        np.testing.assert_array_equal(ctx.raw2rgba(odd, fmt, wq, hq - 1, 0), port.quad2rgba(port.raw2quad(odd, fmt, wq, hq - 1), fmt, 0))
    rgba = rng.integers(0, 256, (n, hq, wq, 4), dtype=np.uint8)
    f32 = (rng.standard_normal((n, hq, wq)) * 150).astype(np.float32)
    f32[0, 0, :3] = [np.nan, np.inf, -np.inf]
    got = ctx.nv12_batch("rgba", rgba, wq, hq, stride=stride)
    for i in range(n):
        np.testing.assert_array_equal(got[i, :used], port.rgba2nv12(rgba[i])[:used])
        assert not got[i, used:].any()
    got = ctx.nv12_batch("f32", f32, wq, hq, stride=stride)
    for i in range(n):
        np.testing.assert_array_equal(got[i, :used], port.f2nv12(f32[i])[:used])
