"""GPU parity of the fused detection path (vp_detect_host / vp_detect_batch_device) against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import common
import oracle as O
from vpb200 import lib

pytestmark = pytest.mark.gpu

CASES = [
    dict(wq=96, hq=64, fmt=0),
    dict(wq=102, hq=66, fmt=0, k2=0.12, tilt=0.2),
    dict(wq=96, hq=64, fmt=1, k2=-0.08, tilt=-0.1, sample_mode=1),
    dict(wq=128, hq=96, fmt=2, k2=0.05, sample_mode=2),
    dict(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.3, n_robots=4, n_balls=3, seed=7),
    dict(wq=160, hq=120, fmt=0, frame="noise", seed=3, thr=15.0),
]


def check_frame(got, i, want):
    np.testing.assert_array_equal(got["counter"][i], want["counter"])
    common.assert_matches_equal(got["matches"][i], want["matches"])


@pytest.mark.parametrize("kw", CASES)
def test_detect_single_frame(ctx, port, kw):
    p, raw, _ = common.make_case(**kw)
    want = port.detect(raw, p)
    got = ctx.detect(raw, common.to_vp(p))
    np.testing.assert_array_equal(got["flat"], want["flat"])
    np.testing.assert_array_equal(got["grad"], want["grad"])
    common.assert_float_images_equal(got["circ"], want["circ"])
    check_frame(got, 0, want)
    assert want["max_abs_sat"] < 2 ** 24 and got["sat_fallbacks"] == 0


@pytest.mark.parametrize("scale_mul,dx,dy", [(2.7, 0.0, 0.0), (0.37, 40.0, -25.0), (1.0, -900.0, 300.0), (1.0, 150.0, -700.0), (1.6, 5000.0, 0.0)])
def test_detect_unusual_geometry(ctx, port, scale_mul, dx, dy):
    """Flat tiles whose source footprint does not fit the staged tile (coarse scale), is tiny (fine scale), or lies
    partly/entirely outside the sensor (shifted extent: CLAMP_TO_EDGE on whole tiles)."""
    p, raw, _ = common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.2, n_robots=3, n_balls=2, seed=11)
    p.field_scale *= scale_mul
    p.off_x += dx
    p.off_y += dy
    want = port.detect(raw, p)
    got = ctx.detect(raw, common.to_vp(p))
    np.testing.assert_array_equal(got["flat"], want["flat"])
    np.testing.assert_array_equal(got["grad"], want["grad"])
    common.assert_float_images_equal(got["circ"], want["circ"])
    check_frame(got, 0, want)
    ctx.set_staged_reproject(False)  # the direct-gather kernel must agree bit for bit
    try:
        direct = ctx.detect(raw, common.to_vp(p))
    finally:
        ctx.set_staged_reproject(2)
    np.testing.assert_array_equal(direct["flat"], want["flat"])


def test_detect_degenerate_camera_gives_nonfinite_coordinates(ctx, port):
    """A camera in the field plane (rz == 0 for every pixel): coordinates are inf/NaN, every tap clamps or reads NaN -> 0."""
    p, raw, _ = common.make_case(wq=96, hq=64)
    p.model.c[2] = p.max_robot_height  # vz = 0 -> division by zero in field2image
    want = port.detect(raw, p)
    got = ctx.detect(raw, common.to_vp(p))
    np.testing.assert_array_equal(got["flat"], want["flat"])
    check_frame(got, 0, want)


@pytest.mark.parametrize("switch", ["staged_reproject", "fused_gradcirc"])
@pytest.mark.parametrize("kw", [dict(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.3, n_robots=4, n_balls=3, seed=7), dict(wq=102, hq=66, fmt=1, k2=0.12, tilt=0.2),
                                dict(wq=160, hq=120, fmt=0, frame="noise", seed=3, max_blobs=64)])
def test_alternative_kernels_agree_with_the_oracle(ctx, port, kw, switch):
    """Every A/B switch selects a different kernel for the same stage (direct-gather vs frame-invariant-hoisted reprojection; gradient + row
    sums + streaming circularity vs the fused gradient + circularity kernel): all settings must give the oracle's bits."""
    p, raw, _ = common.make_case(**kw)
    want = port.detect(raw, p)
    setter = getattr(ctx, "set_" + switch)
    values = {"staged_reproject": (0, 2), "fused_gradcirc": (0, 2, 1)}[switch]  # default last
    default = values[-1]
    try:
        for value in values:
            setter(value)
            got = ctx.detect(np.stack([raw, raw, raw]), common.to_vp(p))
            np.testing.assert_array_equal(got["flat"], want["flat"])
            np.testing.assert_array_equal(got["grad"], want["grad"])
            common.assert_float_images_equal(got["circ"], want["circ"])
            for i in range(3):
                check_frame(got, i, want)
    finally:
        setter(default)


def test_sat_fallback_through_every_flow(ctx, port):
    """The >2^22 fallback (sequential-order SAT, literal 16-tap circularity) behind the fused gradient + circularity kernel, behind the
    row-sum flow, and -- radius 13 -- on the materialised-SAT path of radii outside the specialised range."""
    p, _, _ = common.make_case(wq=256, hq=256, scale_mm=4.0)
    h, w = 2 * p.hq, 2 * p.wq
    yy, xx = np.mgrid[0:h, 0:w]
    raw = (((xx + yy) // 6) % 2 * 255).astype(np.uint8).reshape(-1)
    want = port.detect(raw, p)
    try:
        for gc in (0, 2):  # 2: the fused kernel also on this lone frame (latency path: the bound is checked next to the record kernel)
            ctx.set_fused_gradcirc(gc)
            got = ctx.detect(raw, common.to_vp(p))
            assert got["sat_fallbacks"] == 1
            common.assert_float_images_equal(got["circ"], want["circ"])
            check_frame(got, 0, want)
    finally:
        ctx.set_fused_gradcirc(1)
    p.circle_radius = 13
    want = port.detect(raw, p)
    got = common.detect_device(ctx, [raw, raw], common.to_vp(p), images=("circ",))
    assert got["plan"]["circ"] == 0 and got["sat_fallbacks"] == 2
    common.assert_frame_equal(got, 1, want, images=("circ",))


def test_detect_batch_of_distinct_frames(ctx, port):
    """11 frames (not a multiple of the upload chunk or the launch group), each with its own noise seed."""
    frames, wants = [], []
    for s in range(11):
        p, raw, _ = common.make_case(wq=128, hq=80, seed=s, n_robots=2, n_balls=2)
        frames.append(raw)
        wants.append(port.detect(raw, p, want_images=False))
    got = ctx.detect(np.stack(frames), common.to_vp(p), want_images=False)
    for i, w in enumerate(wants):
        check_frame(got, i, w)


@pytest.mark.parametrize("chunk", [2, 4, 16])
@pytest.mark.parametrize("kw", [dict(wq=128, hq=80, n_robots=2, n_balls=2), dict(wq=102, hq=66, fmt=1, k2=0.12, tilt=0.2, n_robots=2, n_balls=1)])
def test_hoisted_reprojection_over_frame_chunks(ctx, port, chunk, kw):
    """The hoisted reprojection keeps a tile's weights in registers over `chunk` frames: 7 distinct frames in one launch
    group through chunk sizes that divide, do not divide and exceed the group; full and partial tiles, 16-byte-aligned
    and unaligned raw rows; every frame's flat image compared."""
    frames, wants = [], []
    for s in range(7):
        p, raw, _ = common.make_case(seed=20 + s, **kw)
        frames.append(raw)
        wants.append(port.detect(raw, p))
    vp = common.to_vp(p)
    n, nf, rb = len(frames), p.wf * p.hf, frames[0].size
    bufs = dict(raw=ctx.buffer(n * rb, np.stack(frames)), flat=ctx.buffer(n * nf * 4), grad=ctx.buffer(n * nf * 4), circ=ctx.buffer(n * nf * 4),
                m=ctx.buffer(n * vp.max_blobs * 22), c=ctx.buffer(n * 12))
    ctx.set_hoist_chunk(chunk)
    ctx.set_lanes(1)
    ctx.set_group(n)
    try:
        ctx.detect_batch_device(bufs["raw"].device_ptr, n, vp, bufs["flat"].device_ptr, bufs["grad"].device_ptr, bufs["circ"].device_ptr,
                                bufs["m"].device_ptr, bufs["c"].device_ptr)
        flat = bufs["flat"].read(np.uint8).reshape(n, p.hf, p.wf, 4)
        counter = bufs["c"].read(np.int32).reshape(n, 3)
        plan = ctx.last_plan()
    finally:
        ctx.set_hoist_chunk(0)
        ctx.set_group(0)
        ctx.set_lanes(3)
    assert plan["chunk"] == chunk and plan["group"] == n and plan["reproject"] == (4 if chunk >= 4 else 2)
    for i, w in enumerate(wants):
        np.testing.assert_array_equal(flat[i], w["flat"])
        np.testing.assert_array_equal(counter[i], w["counter"])
    for b in bufs.values():
        b.release()


@pytest.mark.parametrize("strips", [1, 2, 5, 16])
@pytest.mark.parametrize("kw", [dict(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.3, n_robots=4, n_balls=3, seed=7),
                                dict(wq=102, hq=66, fmt=1, k2=0.12, tilt=-0.25, n_robots=2, n_balls=1),
                                dict(wq=160, hq=420, fmt=0, frame="noise", seed=3, thr=15.0)])
def test_lone_frame_pipelined_with_its_upload(ctx, port, kw, strips):
    """Latency path: the frame arrives in `strips` chunks of raw rows and reprojection, gradient + row sums and the
    circularity segments run strip by strip behind them.  Every chunking gives the results of the whole-frame pass."""
    p, raw, _ = common.make_case(**kw)
    want = port.detect(raw, p)
    ctx.set_strips(strips)
    try:
        got = ctx.detect(raw, common.to_vp(p), pinned=True)
    finally:
        ctx.set_strips(2)
    np.testing.assert_array_equal(got["flat"], want["flat"])
    np.testing.assert_array_equal(got["grad"], want["grad"])
    common.assert_float_images_equal(got["circ"], want["circ"])
    check_frame(got, 0, want)
    assert got["sat_fallbacks"] == 0


@pytest.mark.parametrize("strips", [3, 16])
def test_lone_frame_in_strips_with_sat_fallback_and_rolled_camera(ctx, port, strips):
    """(a) a frame whose row sums leave the exactness bound in a LATE strip, after earlier strips have published blobs:
    everything published is forgotten and the frame is redone in sequential order; (b) a camera rolled by 180 degrees:
    the first flat rows need the last raw rows, the schedule degenerates to upload-then-compute."""
    p, clean, _ = common.make_case(wq=1600, hq=96, scale_mm=4.0, n_robots=3, n_balls=3, seed=3)
    h, w = 2 * p.hq, 2 * p.wq
    yy, xx = np.mgrid[0:h, 0:w]
    stripes = (((xx + yy) // 6) % 2 * 255).astype(np.uint8)
    late = clean.reshape(h, w).copy()
    late[h - 40:, :3000] = stripes[h - 40:, :3000]
    ctx.set_strips(strips)
    try:
        want = port.detect(late.reshape(-1), p)
        got = ctx.detect(late.reshape(-1), common.to_vp(p), pinned=True)
        assert got["sat_fallbacks"] == 1
        np.testing.assert_array_equal(got["grad"], want["grad"])
        common.assert_float_images_equal(got["circ"], want["circ"])
        check_frame(got, 0, want)

        p2, raw2, _ = common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.2, n_robots=3, n_balls=2, seed=11)
        r = np.array(p2.model.r, np.float32).reshape(3, 3)
        r[0] *= -1.0  # image x and y axes reversed: a roll of 180 degrees about the optical axis
        r[1] *= -1.0
        for i, v in enumerate(r.reshape(-1)):
            p2.model.r[i] = float(v)
        want2 = port.detect(raw2, p2)
        got2 = ctx.detect(raw2, common.to_vp(p2), pinned=True)
        np.testing.assert_array_equal(got2["flat"], want2["flat"])
        check_frame(got2, 0, want2)
    finally:
        ctx.set_strips(2)


@pytest.mark.parametrize("strips", [1, 3, 6])
def test_lone_frames_replayed_as_a_graph_from_a_pinned_ring(ctx, port, strips):
    """A camera's ring of pinned buffers: the first call launches directly, the second is captured into a CUDA graph,
    the following ones replay it with the upload nodes repointed.  Distinct frames (one of them leaving the exactness
    bound, which the host has to redo after a replay) must each give their own oracle result."""
    p, f0, _ = common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.3, n_robots=4, n_balls=3, seed=7)
    h, w = 2 * p.hq, 2 * p.wq
    yy, xx = np.mgrid[0:h, 0:w]
    stripes = (((xx + yy) // 6) % 2 * 255).astype(np.uint8).reshape(-1)
    frames = [f0, common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.3, n_robots=2, n_balls=5, seed=8)[1],
              common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.3, frame="noise", seed=9)[1], stripes]
    wants = [port.detect(f, p) for f in frames]
    assert wants[3]["max_abs_sat"] > 2 ** 24
    vp = common.to_vp(p)
    rb = vp.raw_frame_bytes()
    ring = lib.PinnedArray((len(frames), rb), np.uint8)
    pm = lib.PinnedArray((p.max_blobs,), lib.MATCH_DTYPE)
    pc = lib.PinnedArray((1, 3), np.int32)
    ctx.set_strips(strips)
    launches = []
    try:
        for it in range(3):
            for i, f in enumerate(frames):
                ring.array[i] = f
                pm.array[:] = np.zeros((), lib.MATCH_DTYPE)
                n0 = ctx.launch_count()
                ctx.detect_host_into(ring.ptr.value + i * rb, 1, vp, pm.ptr.value, pc.ptr.value)
                launches.append(ctx.launch_count() - n0)
                np.testing.assert_array_equal(pc.array[0], wants[i]["counter"])
                common.assert_matches_equal(pm.array[: min(int(pc.array[0, 0]), p.max_blobs)].copy(), wants[i]["matches"])
                assert ctx.sat_fallbacks() == (1 if i == 3 else 0)
        # the replays count the kernels they launch like direct calls do (flagged frame: + the redo)
        assert launches[8] == launches[4] == launches[5] and launches[11] == launches[7] > launches[4]
        # same thing with the graph switched off
        ctx.set_latency_graph(False)
        n0 = ctx.launch_count()
        ctx.detect_host_into(ring.ptr.value, 1, vp, pm.ptr.value, pc.ptr.value)
        assert ctx.launch_count() - n0 == launches[4]
        np.testing.assert_array_equal(pc.array[0], wants[0]["counter"])
        common.assert_matches_equal(pm.array[: int(pc.array[0, 0])].copy(), wants[0]["matches"])
    finally:
        ctx.set_latency_graph(True)
        ctx.set_strips(2)
        ring.free(); pm.free(); pc.free()


def test_detect_blob_overflow_keeps_first_in_raster_order(ctx, port):
    p, raw, _ = common.make_case(wq=160, hq=120, frame="noise", seed=9, max_blobs=50)
    want = port.detect(raw, p)
    got = ctx.detect(raw, common.to_vp(p))
    assert want["counter"][0] > 50 and len(want["matches"]) == 50
    check_frame(got, 0, want)


def test_sat_beyond_2p24_falls_back_to_sequential_order(ctx, port):
    """A frame whose SAT leaves the exact-integer range of fp32: the reference's sequential rounding is reproduced."""
    p, _, _ = common.make_case(wq=256, hq=256, scale_mm=4.0)
    # vertical bars shifted by one row per column -> gx*gy has one sign everywhere -> |SAT| grows monotonically
    h, w = 2 * p.hq, 2 * p.wq
    yy, xx = np.mgrid[0:h, 0:w]
    raw = (((xx + yy) // 6) % 2 * 255).astype(np.uint8).reshape(-1)
    want = port.detect(raw, p)
    assert want["max_abs_sat"] > 2 ** 24
    got = ctx.detect(raw, common.to_vp(p))
    assert got["sat_fallbacks"] == 1
    np.testing.assert_array_equal(got["grad"], want["grad"])
    common.assert_float_images_equal(got["circ"], want["circ"])
    check_frame(got, 0, want)


@pytest.mark.parametrize("gc", [1, 0])
def test_flagged_and_clean_frames_in_one_batch(ctx, port, gc):
    """Frames that leave the exactness bound -- one through its row sums (wide stripes), one only through the summed-area
    table (found after the fast pass when there is no SAT) -- between clean frames of the same batch: the flagged ones
    are redone in sequential order, the clean ones keep the results of the fast pass."""
    p, clean, _ = common.make_case(wq=1600, hq=48, scale_mm=4.0, n_robots=3, n_balls=3, seed=3)
    h, w = 2 * p.hq, 2 * p.wq
    yy, xx = np.mgrid[0:h, 0:w]
    stripes = (((xx + yy) // 6) % 2 * 255).astype(np.uint8)
    wide = clean.reshape(h, w).copy()
    wide[:, :3000] = stripes[:, :3000]      # row sums beyond 2^22, blobs to the right of the stripes
    narrow = clean.reshape(h, w).copy()
    narrow[:, :500] = stripes[:, :500]      # row sums stay small, the SAT does not
    frames = [clean, wide.reshape(-1), clean, narrow.reshape(-1), clean]
    wants = [port.detect(f, p) for f in frames]
    rs = [np.abs(np.cumsum(wt["grad"].astype(np.float64), axis=1)).max() for wt in wants]
    assert rs[1] >= 2 ** 22 and rs[3] < 2 ** 22 <= wants[3]["max_abs_sat"] and wants[0]["max_abs_sat"] < 2 ** 21
    vp = common.to_vp(p)
    n, nf, rb = len(frames), p.wf * p.hf, frames[0].size
    bufs = dict(raw=ctx.buffer(n * rb, np.stack(frames)), flat=ctx.buffer(n * nf * 4), grad=ctx.buffer(n * nf * 4), circ=ctx.buffer(n * nf * 4),
                m=ctx.buffer(n * vp.max_blobs * 22), c=ctx.buffer(n * 12))
    ctx.set_fused_gradcirc(gc)
    try:
        ctx.detect_batch_device(bufs["raw"].device_ptr, n, vp, bufs["flat"].device_ptr, bufs["grad"].device_ptr, bufs["circ"].device_ptr,
                                bufs["m"].device_ptr, bufs["c"].device_ptr)
        assert ctx.sat_fallbacks() == 2
        circ = bufs["circ"].read(np.float32).reshape(n, p.hf, p.wf)
        counter = bufs["c"].read(np.int32).reshape(n, 3)
        m = bufs["m"].read(np.uint8).reshape(n, vp.max_blobs, 22)
    finally:
        ctx.set_fused_gradcirc(1)
    for i, wt in enumerate(wants):
        common.assert_float_images_equal(circ[i], wt["circ"])
        np.testing.assert_array_equal(counter[i], wt["counter"])
        k = min(int(counter[i, 0]), vp.max_blobs)
        assert k > 0
        common.assert_matches_equal(m[i, :k].copy().view(lib.MATCH_DTYPE).reshape(-1), wt["matches"])
    for b in bufs.values():
        b.release()


def test_detect_device_pointer_api_matches_host_api(ctx, port):
    p, raw, _ = common.make_case(wq=128, hq=96, n_robots=2, seed=4)
    vp = common.to_vp(p)
    n, nf, rb = 5, p.wf * p.hf, raw.size
    raws = np.stack([np.roll(raw, 2 * 2 * p.wq * k) for k in range(n)])  # shift by whole quad rows
    host = ctx.detect(raws, vp, want_images=False)
    bufs = dict(raw=ctx.buffer(n * rb, raws), flat=ctx.buffer(n * nf * 4), grad=ctx.buffer(n * nf * 4), circ=ctx.buffer(n * nf * 4),
                m=ctx.buffer(n * vp.max_blobs * 22), c=ctx.buffer(n * 12))
    for group in (1, 2, 0):
        ctx.set_group(group)
        ctx.detect_batch_device(bufs["raw"].device_ptr, n, vp, bufs["flat"].device_ptr, bufs["grad"].device_ptr, bufs["circ"].device_ptr,
                                bufs["m"].device_ptr, bufs["c"].device_ptr)
        counter = bufs["c"].read(np.int32).reshape(n, 3)
        m = bufs["m"].read(np.uint8).reshape(n, vp.max_blobs, 22)
        np.testing.assert_array_equal(counter, host["counter"])
        for i in range(n):
            k = int(counter[i, 0])
            np.testing.assert_array_equal(m[i, :k], common.match_bytes(host["matches"][i]))
    for i in range(n):
        check_frame(host, i, port.detect(raws[i], p, want_images=False))
    for b in bufs.values():
        b.release()


def test_full_size_headline_config(ctx, port):
    """BASELINE config 2: one 2448x2048 BayerRG8 frame, full detection, bit-exact against the oracle."""
    from vpb200 import geometry as G, synth as S
    wq, hq = 1224, 1024
    cam = G.default_camera(wq, hq, k2=0.0)
    persp = G.Perspective(cam)
    persp.geometry_check(wq, hq, 180.0)
    lp = G.launch_params(persp, 0, wq, hq)
    scene = S.random_scene(persp.visible_field_extent, 16, 4, seed=1)
    raw = S.render_raw(scene, cam, 2 * wq, 2 * hq, seed=1).reshape(-1)
    p = common.to_vpo(lp)
    want = port.detect(raw, p)
    got = ctx.detect(raw, common.to_vp(p))
    np.testing.assert_array_equal(got["flat"], want["flat"])
    np.testing.assert_array_equal(got["grad"], want["grad"])
    common.assert_float_images_equal(got["circ"], want["circ"])
    check_frame(got, 0, want)
    assert len(want["matches"]) >= 16 * 5  # every pattern blob is a peak


def test_full_size_with_host_derived_parameters(ctx, port):
    """Calibration -> vp_camera_model_from_calib -> vp_geometry_check -> vp_geometry_params -> fused detection: the launch
    scalars come from the C++ host derivation (reference summation order: 3.987 mm/px, flat 1208 x 1010), the frame is the
    headline 2448x2048 scene; bit-exact against the oracle run with the same scalars."""
    from vpb200 import geometry as G, synth as S
    wq, hq = 1224, 1024
    cam = G.default_camera(wq, hq, k2=0.03)
    w, x, y, z = cam.quat_wxyz
    t = cam.f2i() @ (-np.asarray(cam.pos, np.float32))
    calib = lib.CameraCalib(wq, hq, cam.focal_length, cam.principal_point[0], cam.principal_point[1], cam.distortion_k2, x, y, z, w,
                            float(t[0]), float(t[1]), float(t[2]))
    f = G.FieldSize()
    hp = lib.HostPerspective(calib, lib.FieldSizeC(f.field_length, f.field_width, f.boundary_width, f.boundary_width_goal_line, f.ball_radius))
    hp.geometry_check(wq, hq, 180.0)
    vp = hp.params(0, wq, hq)
    assert vp.wf % 2 == 0 and vp.hf % 2 == 0 and (vp.wf, vp.hf) != (wq, hq)
    scene = S.random_scene(hp.visible_field_extent, 16, 4, seed=2)
    raw = S.render_raw(scene, cam, 2 * wq, 2 * hq, seed=2).reshape(-1)
    po = O.Params()
    assert C.sizeof(po) == C.sizeof(vp)
    C.memmove(C.byref(po), C.byref(vp), C.sizeof(vp))
    want = port.detect(raw, po)
    got = ctx.detect(raw, vp)
    np.testing.assert_array_equal(got["flat"], want["flat"])
    np.testing.assert_array_equal(got["grad"], want["grad"])
    common.assert_float_images_equal(got["circ"], want["circ"])
    check_frame(got, 0, want)
    assert want["counter"][0] >= 60


def test_full_size_batch_through_the_four_frame_reprojection(ctx, port):
    """Six distinct 2448x2048 frames in one device-resident batch: the automatic chunk is 4, so the byte-packed four-frame
    reprojection runs a full quad and a partial one (two frames); every frame's flat image and blob list against the oracle."""
    from vpb200 import geometry as G, synth as S
    wq, hq = 1224, 1024
    cam = G.default_camera(wq, hq, k2=0.05)
    persp = G.Perspective(cam)
    persp.geometry_check(wq, hq, 180.0)
    lp = G.launch_params(persp, 0, wq, hq)
    p = common.to_vpo(lp)
    vp = common.to_vp(p)
    scene = S.random_scene(persp.visible_field_extent, 12, 3, seed=9)
    clean = S.render_rgb(scene, cam, 2 * wq, 2 * hq)
    frames = [S.render_raw(scene, cam, 2 * wq, 2 * hq, S.FMT_RGGB, seed=100 + i, clean_rgb=clean).reshape(-1) for i in range(6)]
    n, nf, rb = len(frames), p.wf * p.hf, frames[0].size
    bufs = dict(raw=ctx.buffer(n * rb, np.stack(frames)), flat=ctx.buffer(n * nf * 4), grad=ctx.buffer(n * nf * 4), circ=ctx.buffer(n * nf * 4),
                m=ctx.buffer(n * vp.max_blobs * 22), c=ctx.buffer(n * 12))
    ctx.detect_batch_device(bufs["raw"].device_ptr, n, vp, bufs["flat"].device_ptr, bufs["grad"].device_ptr, bufs["circ"].device_ptr,
                            bufs["m"].device_ptr, bufs["c"].device_ptr)
    flat = bufs["flat"].read(np.uint8).reshape(n, p.hf, p.wf, 4)
    circ = bufs["circ"].read(np.float32).reshape(n, p.hf, p.wf)
    counter = bufs["c"].read(np.int32).reshape(n, 3)
    m = bufs["m"].read(np.uint8).reshape(n, vp.max_blobs, 22)
    for i in (0, 3, 4, 5):
        want = port.detect(frames[i], p)
        np.testing.assert_array_equal(flat[i], want["flat"])
        common.assert_float_images_equal(circ[i], want["circ"])
        np.testing.assert_array_equal(counter[i], want["counter"])
        k = min(int(counter[i, 0]), vp.max_blobs)
        common.assert_matches_equal(m[i, :k].copy().view(lib.MATCH_DTYPE).reshape(-1), want["matches"])
    for b in bufs.values():
        b.release()


def test_blobs_to_field_records_and_cell_lists(ctx, port):
    """the CPU loop of main.cpp:297-325 (CLMatch -> Match in field millimetres, spatial index) on the device for
    a batch: records bit-exact against flat2field on the host, the cell lists against a stable sort by (cell, index);
    frames with zero blobs, with an overflowing counter and with NaN positions (plateau peaks) included."""
    frames, lists, counters = [], [], []
    cases = [dict(wq=160, hq=120, n_robots=3, n_balls=3, seed=2), dict(wq=160, hq=120, frame="noise", seed=9, max_blobs=300),
             dict(wq=160, hq=120, n_robots=0, n_balls=0, seed=4, thr=1e9)]
    max_blobs = 300
    for kw in cases:
        kw = dict(kw, max_blobs=max_blobs)
        p, raw, _ = common.make_case(**kw)
        w = port.detect(raw, p, want_images=False)
        rec = np.zeros(max_blobs, lib.MATCH_DTYPE)
        rec[:len(w["matches"])] = w["matches"]
        lists.append(rec)
        counters.append(w["counter"])
    plateau = np.zeros(max_blobs, lib.MATCH_DTYPE)   # hand-made list: NaN offsets, positions outside the grid
    plateau["x"][:4] = [np.nan, -50.0, 1e6, 17.25]
    plateau["y"][:4] = [3.0, np.nan, -1e6, 40.5]
    lists.append(plateau)
    counters.append(np.array([4, 0, 0], np.int32))
    matches, counters = np.stack(lists), np.stack(counters)
    assert counters[1, 0] > max_blobs and counters[2, 0] == 0
    scale, ox, oy, cell = np.float32(p.field_scale), np.float32(p.off_x), np.float32(p.off_y), np.float32(150.0)
    cx, cy = 6, 5
    rec, order, cs = ctx.blobs_to_field(matches, counters, max_blobs, float(scale), float(ox), float(oy), float(cell), cx, cy)
    for f in range(len(counters)):
        n = min(int(counters[f, 0]), max_blobs)
        m = matches[f, :n]
        pos = np.stack([m["x"] * scale + ox, m["y"] * scale + oy], -1).astype(np.float32)       # Perspective.cpp:127-129
        common.assert_float_images_equal(rec[f, :n]["pos"], pos)        # bit-exact; NaN by class (payloads are platform-specific)
        np.testing.assert_array_equal(rec[f, :n]["color"], m["color"].astype(np.int32))
        np.testing.assert_array_equal(rec[f, :n]["center"], m["center"].astype(np.int32))
        np.testing.assert_array_equal(rec[f, :n]["circ"].view(np.uint32), m["circ"].view(np.uint32))
        np.testing.assert_array_equal(rec[f, :n]["score"].view(np.uint32), m["score"].view(np.uint32))
        with np.errstate(invalid="ignore"):
            ix = np.floor((pos[:, 0] - ox) / cell)
            iy = np.floor((pos[:, 1] - oy) / cell)
        ix = np.where(ix >= 0, np.minimum(ix, cx - 1), 0).astype(np.int64)      # NaN compares false -> cell 0
        iy = np.where(iy >= 0, np.minimum(iy, cy - 1), 0).astype(np.int64)
        cells = iy * cx + ix
        want_order = np.argsort(cells, kind="stable")
        np.testing.assert_array_equal(order[f, :n], want_order)
        want_cs = np.searchsorted(cells[want_order], np.arange(cx * cy + 1), side="left")
        np.testing.assert_array_equal(cs[f], want_cs)


@pytest.mark.parametrize("scale_mul,dx,dy", [(2.7, 0.0, 0.0), (0.37, 40.0, -25.0), (1.0, -900.0, 300.0), (1.6, 5000.0, 0.0)])
def test_four_frame_reprojection_on_unusual_geometry(ctx, port, scale_mul, dx, dy):
    """The byte-packed four-frame kernel on tiles whose footprint does not fit (direct-gather fallback inside the kernel), is
    tiny, or lies partly/entirely outside the sensor (edge replication in the byte transpose): five frames, chunk 4."""
    frames = []
    for s_ in range(5):
        p, raw, _ = common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.2, n_robots=3, n_balls=2, seed=30 + s_)
        frames.append(raw)
    p.field_scale *= scale_mul
    p.off_x += dx
    p.off_y += dy
    vp = common.to_vp(p)
    n, nf, rb = len(frames), p.wf * p.hf, frames[0].size
    bufs = dict(raw=ctx.buffer(n * rb, np.stack(frames)), flat=ctx.buffer(n * nf * 4), grad=ctx.buffer(n * nf * 4), circ=ctx.buffer(n * nf * 4),
                m=ctx.buffer(n * vp.max_blobs * 22), c=ctx.buffer(n * 12))
    ctx.set_hoist_chunk(4)
    try:
        ctx.detect_batch_device(bufs["raw"].device_ptr, n, vp, bufs["flat"].device_ptr, bufs["grad"].device_ptr, bufs["circ"].device_ptr,
                                bufs["m"].device_ptr, bufs["c"].device_ptr)
        flat = bufs["flat"].read(np.uint8).reshape(n, p.hf, p.wf, 4)
        counter = bufs["c"].read(np.int32).reshape(n, 3)
    finally:
        ctx.set_hoist_chunk(0)
    for i in range(n):
        want = port.detect(frames[i], p)
        np.testing.assert_array_equal(flat[i], want["flat"])
        np.testing.assert_array_equal(counter[i], want["counter"])
    for b in bufs.values():
        b.release()


def test_a_second_context_after_the_first_was_destroyed(port):
    """Contexts come and go within one process (one per camera, src/main.cpp runs one per process but tools do not): destroying
    one must leave no CUDA error behind for the next one's first launch to trip over."""
    p, raw, _ = common.make_case(wq=96, hq=64, seed=2)
    want = port.detect(raw, p, want_images=False)
    for _ in range(2):
        with lib.Context(0) as c:
            got = c.detect(raw, common.to_vp(p), want_images=False)
            check_frame(got, 0, want)
            c.detect(np.stack([raw] * 5), common.to_vp(p), want_images=False)
