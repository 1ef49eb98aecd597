"""CPU: the integer form of quad2nv12.cl / quad2rgba.cl that k_raw2nv12_wide and k_raw2rgba_wide compute on 16-bit lanes.

quad2nv12.cl:36-40 asks the LINEAR sampler for integer positions +-0.25, so every filter weight is 1/4 or 3/4 against the
texel to the left / above and the float blend of the reference is S/16 with an integer S <= 4080.  This file pins, without a
GPU, (a) that the kernel's rounding expression is round-half-to-even of S/16 for EVERY possible S, and (b) that a numpy
restatement of the whole integer pipeline (the arithmetic of the kernel, not its code) reproduces the oracle's bytes -- and
the reference kernels' own, where they are built -- on random frames in both Bayer orders."""
import numpy as np
import pytest


def rte16(s):
    """(S + 7 + ((S >> 4) & 1)) >> 4 -- rte16_lanes() of kernels.cuh on one lane."""
    return (s + 7 + ((s >> 4) & 1)) >> 4


def test_rounding_expression_is_round_half_to_even_for_every_numerator():
    s = np.arange(0, 4081, dtype=np.int64)
    as_float = (s.astype(np.float32) / np.float32(16.0)).astype(np.float32)   # exact: S <= 4080 has 12 bits
    assert np.array_equal(as_float.astype(np.float64) * 16.0, s.astype(np.float64))
    assert np.array_equal(rte16(s), np.rint(as_float).astype(np.int64))        # np.rint rounds half to even like __float2uint_rn


def integer_demosaic(raw, fmt, wq, hq):
    """r, g, b (hq x wq, int) of quad2nv12.cl's demosaic at default sampling, integer arithmetic only."""
    img = raw.reshape(2 * hq, 2 * wq).astype(np.int64)
    planes = [img[0::2, 0::2], img[0::2, 1::2], img[1::2, 0::2], img[1::2, 1::2]]  # raw2quad.cl:31-37

    def left(a):   # texel x-1, CLAMP_TO_EDGE
        return np.concatenate([a[:, :1], a[:, :-1]], axis=1)

    def above(a):  # texel y-1, CLAMP_TO_EDGE
        return np.concatenate([a[:1, :], a[:-1, :]], axis=0)

    v = []
    for c, t in enumerate(planes):
        # plane 0 is tapped at (+.25, +.25), 1 at (-.25, +.25), 2 at (+.25, -.25), 3 at (-.25, -.25) (quad2nv12.cl:36-40)
        h = 3 * t + left(t) if c % 2 == 0 else 3 * left(t) + t        # +0.25: left 1/4, own 3/4;  -0.25: left 3/4, own 1/4
        s = 3 * h + above(h) if c < 2 else 3 * above(h) + h           # +0.25: above 1/4, own 3/4; -0.25: above 3/4, own 1/4
        assert s.max() <= 4080
        v.append(rte16(s))
    if fmt == 0:   # RGGB, quad2nv12.cl:41-43
        return v[0], v[1] // 2 + v[2] // 2, v[3]
    return v[1], v[0] // 2 + v[3] // 2, v[2]   # GRBG


def integer_nv12(raw, fmt, wq, hq):
    r, g, b = integer_demosaic(raw, fmt, wq, hq)
    y = np.minimum((66 * r + 129 * g + 25 * b) // 256 + 16, 255)      # rgba2nv12.cl:27 (unsigned)
    rr, gg, bb = r[1::2, 1::2], g[1::2, 1::2], b[1::2, 1::2]          # last writer of the racing UV stores: bottom-right pixel
    tz = lambda a: np.where(a >= 0, a // 256, -((-a) // 256))         # C division truncates toward zero
    u = np.clip(tz(-38 * rr - 74 * gg + 112 * bb) + 128, 0, 255)
    v = np.clip(tz(112 * rr - 94 * gg - 18 * bb) + 128, 0, 255)
    uv = np.stack([u, v], axis=-1).reshape(hq // 2, wq)
    return np.concatenate([y.reshape(-1), uv.reshape(-1)]).astype(np.uint8)


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("wq,hq", [(72, 46), (8, 2), (40, 30)])
def test_integer_pipeline_reproduces_the_oracle(port, fmt, wq, hq):
    rng = np.random.default_rng(100 * fmt + wq)
    for k in range(3):
        raw = rng.integers(0, 256, 4 * wq * hq, dtype=np.uint8)
        if k == 1:
            raw[: 4 * wq] = np.tile(np.array([0, 255, 255, 0], np.uint8), wq)   # extremes side by side
        ch = port.raw2quad(raw, fmt, wq, hq)
        n = wq * hq * 3 // 2
        np.testing.assert_array_equal(integer_nv12(raw, fmt, wq, hq), port.quad2nv12(ch, fmt, 0)[:n])
        r, g, b = integer_demosaic(raw, fmt, wq, hq)
        rgba = port.quad2rgba(ch, fmt, 0).reshape(hq, wq, 4)
        np.testing.assert_array_equal(np.stack([r, g, b], -1).astype(np.uint8), rgba[..., :3])
        assert (rgba[..., 3] == 255).all()


def test_integer_pipeline_reproduces_the_reference_kernels(clref):
    rng = np.random.default_rng(7)
    wq, hq = 56, 34
    for fmt in (0, 1):
        raw = rng.integers(0, 256, 4 * wq * hq, dtype=np.uint8)
        ch = clref.raw2quad(raw, fmt, wq, hq)
        np.testing.assert_array_equal(integer_nv12(raw, fmt, wq, hq), clref.quad2nv12(ch, fmt, 0)[: wq * hq * 3 // 2])
