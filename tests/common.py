"""Shared test scenarios: seeded synthetic cameras/frames and Params conversion (test infrastructure)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

import oracle as O
from vpb200 import geometry as G, synth as S


def make_case(wq=96, hq=64, fmt=0, k2=0.0, seed=1, scale_mm=4.0, tilt=0.0, n_robots=1, n_balls=1, sample_mode=0,
              max_blobs=2000, thr=15.0, frame="scene"):
    """A small camera over the field centre whose flat pixels are ~scale_mm wide, plus one raw frame.

    Returns (vpo_params, raw frame bytes (1-D uint8), LaunchParams)."""
    height = 180.0 + scale_mm * wq
    q = (0.0, -1.0, 0.0, 0.0)
    if tilt:
        # rotate the default top-down orientation by `tilt` rad about the x axis
        c, s = math.cos(tilt / 2), math.sin(tilt / 2)
        # q = q_tilt * q0 with q0 = (0,-1,0,0), q_tilt = (c, s, 0, 0)
        q = (c * 0.0 - s * -1.0, c * -1.0 + s * 0.0, 0.0, 0.0)
    cam = G.CameraModel(size=(wq, hq), focal_length=float(wq), principal_point=(wq / 2.0 + 0.3, hq / 2.0 - 0.2),
                        distortion_k2=k2, pos=(10.0, -5.0, height), quat_wxyz=q)
    persp = G.Perspective(cam)
    persp.geometry_check(wq, hq, 180.0)
    lp = G.launch_params(persp, fmt, wq, hq, circ_threshold=thr, max_blobs=max_blobs, sample_mode=sample_mode)
    sw, sh = (wq, hq) if fmt == S.FMT_BGR else (2 * wq, 2 * hq)
    if frame == "noise":
        raw = S.noise_frame(sw, sh, seed, fmt)
    else:
        ext = persp.visible_field_extent
        sc = S.random_scene(ext, n_robots, n_balls, seed)
        raw = S.render_raw(sc, cam, sw, sh, fmt, seed=seed)
    return to_vpo(lp), np.ascontiguousarray(raw).reshape(-1), lp


def to_vpo(lp: G.LaunchParams) -> O.Params:
    p = O.Params()
    p.fmt, p.wq, p.hq, p.wf, p.hf = lp.fmt, lp.wq, lp.hq, lp.wf, lp.hf
    C.memmove(C.byref(p.model), lp.model_bytes, 72)
    p.max_robot_height, p.field_scale, p.off_x, p.off_y = lp.max_robot_height, lp.field_scale, lp.off_x, lp.off_y
    p.grad_offset, p.circle_radius = lp.grad_offset, lp.circle_radius
    p.circ_threshold, p.min_score = lp.circ_threshold, lp.min_score
    p.blob_radius, p.max_blobs, p.sample_mode = lp.blob_radius, lp.max_blobs, lp.sample_mode
    return p


def to_vp(p: O.Params):
    """vpo_params -> vp_params: the two structs share one layout by design."""
    from vpb200 import lib
    assert C.sizeof(lib.Params) == C.sizeof(O.Params)
    q = lib.Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    return q


def match_bytes(m: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(m).view(np.uint8).reshape(-1, 22)


def canon_nan(m: np.ndarray) -> np.ndarray:
    """NaNs compare as a class (their sign/payload is platform-specific: x86 0/0 = 0xFFC00000, sm_100 = 0x7FFFFFFF)."""
    m = np.array(m, copy=True)
    for f in ("x", "y", "circ", "score"):
        v = m[f].copy()
        v[np.isnan(v)] = np.float32(np.nan)
        m[f] = v
    return m


def assert_matches_equal(a: np.ndarray, b: np.ndarray, ordered=True):
    """Bit-exact comparison of two blob lists (all 22 bytes of every record; NaN offsets of plateaus by class)."""
    assert len(a) == len(b), (len(a), len(b))
    a, b = canon_nan(a), canon_nan(b)
    if not ordered:
        a, b = O.canonical(a), O.canonical(b)
    np.testing.assert_array_equal(match_bytes(a), match_bytes(b))


def assert_float_images_equal(a: np.ndarray, b: np.ndarray):
    """Bit-exact for every finite value and for the sign of infinities; NaN positions must coincide."""
    fin = np.isfinite(a) & np.isfinite(b)
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_array_equal(a[fin].view(np.uint32), b[fin].view(np.uint32))
    inf = ~fin & ~np.isnan(a)
    np.testing.assert_array_equal(a[inf], b[inf])


def detect_device(ctx, frames, vp, images=("flat", "grad", "circ")):
    """`frames` (n raw frames) through vp_detect_batch_device with device-resident buffers; returns every frame's images,
    counters and 22-byte records, and what the call launched (ctx.last_plan())."""
    from vpb200 import lib
    frames = np.ascontiguousarray(np.stack(frames), np.uint8)
    n, rb = frames.shape
    nf = vp.wf * vp.hf
    bufs = dict(raw=ctx.buffer(n * rb, frames), flat=ctx.buffer(n * nf * 4), grad=ctx.buffer(n * nf * 4), circ=ctx.buffer(n * nf * 4),
                m=ctx.buffer(n * max(vp.max_blobs, 1) * 22), c=ctx.buffer(n * 12))
    try:
        ctx.detect_batch_device(bufs["raw"].device_ptr, n, vp, bufs["flat"].device_ptr, bufs["grad"].device_ptr, bufs["circ"].device_ptr,
                                bufs["m"].device_ptr, bufs["c"].device_ptr)
        out = dict(plan=ctx.last_plan(), sat_fallbacks=ctx.sat_fallbacks())
        if "flat" in images:
            out["flat"] = bufs["flat"].read(np.uint8).reshape(n, vp.hf, vp.wf, 4)
        if "grad" in images:
            out["grad"] = bufs["grad"].read(np.float32).reshape(n, vp.hf, vp.wf)
        if "circ" in images:
            out["circ"] = bufs["circ"].read(np.float32).reshape(n, vp.hf, vp.wf)
        counter = bufs["c"].read(np.int32).reshape(n, 3)
        m = bufs["m"].read(np.uint8).reshape(n, max(vp.max_blobs, 1), 22)
        out["counter"] = counter
        out["matches"] = [m[i, : min(int(counter[i, 0]), vp.max_blobs)].copy().view(lib.MATCH_DTYPE).reshape(-1) for i in range(n)]
    finally:
        for b in bufs.values():
            b.release()
    return out


def assert_frame_equal(got, i, want, images=("flat", "grad", "circ")):
    """Frame i of a detect_device() result against one oracle result: images bit for bit, counters, records."""
    if "flat" in images:
        np.testing.assert_array_equal(got["flat"][i], want["flat"])
    if "grad" in images:
        np.testing.assert_array_equal(got["grad"][i], want["grad"])
    if "circ" in images:
        assert_float_images_equal(got["circ"][i], want["circ"])
    np.testing.assert_array_equal(got["counter"][i], want["counter"])
    assert_matches_equal(got["matches"][i], want["matches"])
