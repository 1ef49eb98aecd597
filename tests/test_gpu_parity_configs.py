"""GPU parity of the configurations the benchmarks time but round 1 never compared with the oracle: every circularity
radius the fused kernels are instantiated for (plus the generic path), the 4096x3000 and 1920x1200 sensors of BASELINE
configs 5 and 1, a gradient offset of zero on the fused path, GRBG through the four-frame reprojection, and the strongly
distorted, tilted full-size camera of SURVEY 8(d).  Every test states -- and asserts through vp_detect_last_plan --
which specialised kernel it went through."""
import numpy as np
import pytest

import common
from vpb200 import geometry as G, synth as S

pytestmark = pytest.mark.gpu


def full_size_case(sensor_w, sensor_h, k2=0.0, tilt=0.0, fmt=0, n_frames=1, n_robots=12, n_balls=3, seed=1):
    """The benchmark camera (bench.py / tools/sweep_configs.py: top-down at 5 m, f = wq) over `n_frames` noise seeds of one scene."""
    wq, hq = sensor_w // 2, sensor_h // 2
    cam = G.default_camera(wq, hq, k2=k2)
    if tilt:
        import math
        c, s = math.cos(tilt / 2), math.sin(tilt / 2)
        cam = G.CameraModel(size=cam.size, focal_length=cam.focal_length, principal_point=cam.principal_point, distortion_k2=k2, pos=cam.pos,
                            quat_wxyz=(s, -c, 0.0, 0.0))
    persp = G.Perspective(cam)
    persp.geometry_check(wq, hq, 180.0)
    lp = G.launch_params(persp, fmt, wq, hq)
    scene = S.random_scene(persp.visible_field_extent, n_robots, n_balls, seed=seed)
    clean = S.render_rgb(scene, cam, sensor_w, sensor_h)
    frames = [S.render_raw(scene, cam, sensor_w, sensor_h, fmt, seed=100 * seed + i, clean_rgb=clean).reshape(-1) for i in range(n_frames)]
    return common.to_vpo(lp), frames


@pytest.mark.parametrize("flow", ["gradcirc", "rowsums"])
@pytest.mark.parametrize("radius", list(range(1, 14)))
def test_every_circle_radius_single_frame_and_batch(ctx, port, radius, flow):
    """circle_radius 1..12 are template instantiations of the fused gradient + circularity kernel (flow 'gradcirc', the
    default) and of the streaming circularity kernel over row sums ('rowsums'); 13 takes the materialised-SAT path with the
    unfused circle kernel; a lone frame (host API, latency path) and a batch of five distinct frames (device API) each
    against the oracle."""
    frames = []
    for s_ in range(5):
        p, raw, _ = common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.2, n_robots=3, n_balls=2, seed=40 + s_,
                                     thr=0.0 if radius == 1 else 6.0)  # radius 1: empty boxes, every circularity is 0 -> all pixels are plateau peaks
        frames.append(raw)
    p.circle_radius = radius
    p.blob_radius = max(radius - 1, 0)
    wants = [port.detect(f, p) for f in frames]
    assert any(len(w["matches"]) > 0 for w in wants)
    vp = common.to_vp(p)
    ctx.set_fused_gradcirc(2 if flow == "gradcirc" else 0)
    try:
        got1 = ctx.detect(frames[0], vp)
        got = common.detect_device(ctx, frames, vp)
    finally:
        ctx.set_fused_gradcirc(1)
    np.testing.assert_array_equal(got1["flat"], wants[0]["flat"])
    np.testing.assert_array_equal(got1["grad"], wants[0]["grad"])
    common.assert_float_images_equal(got1["circ"], wants[0]["circ"])
    np.testing.assert_array_equal(got1["counter"][0], wants[0]["counter"])
    common.assert_matches_equal(got1["matches"][0], wants[0]["matches"])
    for i, w in enumerate(wants):
        common.assert_frame_equal(got, i, w)
    gc_fits = p.grad_offset <= 4 and 2 * p.grad_offset <= radius + 2  # else the fused kernel hands over to the row-sum flow
    assert got["plan"]["circ"] == (0 if radius == 13 else {"gradcirc": 4 if gc_fits else 3, "rowsums": 3}[flow])


@pytest.mark.parametrize("offset", [0, 1, 3])
def test_gradient_offsets_on_the_fused_path(ctx, port, offset):
    """gradientDot's offset is (int)ceilf(maxBlobRadius/fieldScale)/3 with integer division (Resources.cpp:160): 0 for coarse
    field scales (gx = gy = 0 everywhere, SURVEY 9 gotcha 2), odd values take the unaligned tap loads."""
    frames = [common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.2, n_robots=3, n_balls=2, seed=50 + s_)[1] for s_ in range(3)]
    p = common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.2)[0]
    p.grad_offset = offset
    wants = [port.detect(f, p) for f in frames]
    if offset == 0:
        assert not wants[0]["grad"].any() and wants[0]["counter"][0] == 0
    got = common.detect_device(ctx, frames, common.to_vp(p))
    for i, w in enumerate(wants):
        common.assert_frame_equal(got, i, w)
    got1 = ctx.detect(frames[1], common.to_vp(p))
    np.testing.assert_array_equal(got1["grad"], wants[1]["grad"])
    common.assert_float_images_equal(got1["circ"], wants[1]["circ"])


def test_coarse_field_scale_gives_offset_zero_through_the_host_derivation(ctx, port):
    """field scale >= 12.5 mm/px: ceil(25/s)/3 == 0 comes out of the launch-parameter derivation itself."""
    p, raw, lp = common.make_case(wq=160, hq=120, scale_mm=13.0, n_robots=2, n_balls=2, seed=3)
    assert p.grad_offset == 0 and p.circle_radius == 2
    want = port.detect(raw, p)
    got = common.detect_device(ctx, [raw, raw], common.to_vp(p))
    common.assert_frame_equal(got, 0, want)
    common.assert_frame_equal(got, 1, want)


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("n,chunk", [(5, 4), (8, 8), (7, 16), (6, 2)])
def test_four_frame_reprojection_is_really_reached(ctx, port, fmt, n, chunk):
    """k_reproject_hoist4 (the headline kernel) with RGGB and GRBG frames, frame counts that are and are not multiples of
    four, chunks that divide, do not divide and exceed the group, partial tiles (102x66) and a tilted, distorted camera.  One
    lane and one group, so that the group really holds all n frames; chunk 2 must select the one-frame-per-word kernel."""
    frames = []
    for s_ in range(n):
        p, raw, _ = common.make_case(wq=102 if fmt else 328, hq=66 if fmt else 200, fmt=fmt, k2=0.12, tilt=0.2, n_robots=3, n_balls=2, seed=60 + s_)
        frames.append(raw)
    wants = [port.detect(f, p) for f in frames]
    ctx.set_lanes(1)
    ctx.set_group(n)
    ctx.set_hoist_chunk(chunk)
    try:
        got = common.detect_device(ctx, frames, common.to_vp(p))
    finally:
        ctx.set_hoist_chunk(0)
        ctx.set_group(0)
        ctx.set_lanes(3)
    assert got["plan"]["group"] == n and got["plan"]["chunk"] == chunk and got["plan"]["lanes"] == 1
    assert got["plan"]["reproject"] == (4 if chunk >= 4 else 2)
    for i, w in enumerate(wants):
        common.assert_frame_equal(got, i, w)


@pytest.mark.parametrize("scale_mul,dx,dy", [(2.7, 0.0, 0.0), (0.37, 40.0, -25.0), (1.0, -900.0, 300.0), (1.6, 5000.0, 0.0)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_four_frame_reprojection_on_unusual_geometry_both_bayer_orders(ctx, port, fmt, scale_mul, dx, dy):
    """Tiles whose footprint does not fit the staged capacity (gather fallback inside the kernel), tiny footprints, and tiles
    partly or entirely outside the sensor (edge replication), five frames through the four-frame kernel (one full quad of
    frames and a partial one), both Bayer orders."""
    frames = []
    for s_ in range(5):
        p, raw, _ = common.make_case(wq=320, hq=200, fmt=fmt, k2=0.1, tilt=0.2, n_robots=3, n_balls=2, seed=30 + s_)
        frames.append(raw)
    p.field_scale *= scale_mul
    p.off_x += dx
    p.off_y += dy
    wants = [port.detect(f, p) for f in frames]
    ctx.set_lanes(1)
    ctx.set_group(5)
    ctx.set_hoist_chunk(4)
    try:
        got = common.detect_device(ctx, frames, common.to_vp(p))
    finally:
        ctx.set_hoist_chunk(0)
        ctx.set_group(0)
        ctx.set_lanes(3)
    assert got["plan"]["reproject"] == 4 and got["plan"]["chunk"] == 4 and got["plan"]["group"] == 5
    for i, w in enumerate(wants):
        common.assert_frame_equal(got, i, w)


def test_config5_sensor_4096x3000_radius_9(ctx, port):
    """BASELINE config 5: 4096x3000 BayerRG8, circle radius 9 (2.35 mm per flat pixel): a lone frame through the host API and
    a batch of three through the device API (automatic group/chunk), every image and blob list against the oracle."""
    p, frames = full_size_case(4096, 3000, n_frames=3)
    assert (p.wq, p.hq) == (2048, 1500) and p.circle_radius == 9
    wants = [port.detect(f, p) for f in frames]
    assert wants[0]["counter"][0] >= 12 * 5
    vp = common.to_vp(p)
    got1 = ctx.detect(frames[0], vp)
    np.testing.assert_array_equal(got1["flat"], wants[0]["flat"])
    np.testing.assert_array_equal(got1["grad"], wants[0]["grad"])
    common.assert_float_images_equal(got1["circ"], wants[0]["circ"])
    np.testing.assert_array_equal(got1["counter"][0], wants[0]["counter"])
    common.assert_matches_equal(got1["matches"][0], wants[0]["matches"])
    got = common.detect_device(ctx, frames, vp)
    for i, w in enumerate(wants):
        common.assert_frame_equal(got, i, w)
    assert got["sat_fallbacks"] == 0


def test_config5_four_frame_kernel_at_4096x3000(ctx, port):
    """The same sensor through the kernels the throughput sweep times: one group of 8 frames, chunk 8, four frames per word."""
    p, frames = full_size_case(4096, 3000, n_frames=8, seed=2)
    ctx.set_lanes(1)
    ctx.set_group(8)
    ctx.set_hoist_chunk(8)
    try:
        got = common.detect_device(ctx, frames, common.to_vp(p), images=("flat", "circ"))
    finally:
        ctx.set_hoist_chunk(0)
        ctx.set_group(0)
        ctx.set_lanes(3)
    assert got["plan"]["reproject"] == 4 and got["plan"]["group"] == 8
    for i in (0, 3, 6, 7):
        common.assert_frame_equal(got, i, port.detect(frames[i], p), images=("flat", "circ"))


def test_config1_sensor_1920x1200(ctx, port):
    """BASELINE config 1's sensor (blob_benchmark on 1920x1200 BayerRG8): flat size and radii differ from the headline."""
    p, frames = full_size_case(1920, 1200, n_frames=2, n_robots=8, n_balls=2)
    assert (p.wq, p.hq) == (960, 600)
    wants = [port.detect(f, p) for f in frames]
    got1 = ctx.detect(frames[0], common.to_vp(p))
    np.testing.assert_array_equal(got1["flat"], wants[0]["flat"])
    common.assert_float_images_equal(got1["circ"], wants[0]["circ"])
    got = common.detect_device(ctx, frames, common.to_vp(p))
    for i, w in enumerate(wants):
        common.assert_frame_equal(got, i, w)


def test_full_size_distorted_tilted_camera_through_the_four_frame_kernel(ctx, port):
    """SURVEY 8(d)'s stress camera: k2 = 0.12 and a tilt of 0.2 rad at 2448x2048.  Six frames in one group (a full quad of
    frames and a partial one); tiles whose footprint exceeds the staged capacity take the gather path inside the kernel."""
    p, frames = full_size_case(2448, 2048, k2=0.12, tilt=0.2, n_frames=6, seed=3)
    ctx.set_lanes(1)
    ctx.set_group(6)
    ctx.set_hoist_chunk(8)
    try:
        got = common.detect_device(ctx, frames, common.to_vp(p))
    finally:
        ctx.set_hoist_chunk(0)
        ctx.set_group(0)
        ctx.set_lanes(3)
    assert got["plan"]["reproject"] == 4 and got["plan"]["group"] == 6
    for i in (0, 4, 5):
        common.assert_frame_equal(got, i, port.detect(frames[i], p))
    # the staged kernel really carried this camera: vp_tile_stats counts the tiles whose footprint fits the staged planes
    st = ctx.tile_stats(common.to_vp(p))
    assert st["tiles"] == -(-p.wf // 64) * -(-p.hf // 16) and 0 < st["staged"] <= st["tiles"] and st["max_rows"] <= 24
    assert st["staged"] >= 0.9 * st["tiles"], st


def test_automatic_groups_are_balanced_whole_quads(ctx, port):
    """13 frames on three streams: the library cuts them into groups of the same size rounded up to whole quads of frames (8 + 5,
    not 4 + 4 + 4 + 1), each group through the four-frame reprojection; every frame against the oracle."""
    frames = [common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.2, n_robots=3, n_balls=2, seed=70 + s_)[1] for s_ in range(13)]
    p = common.make_case(wq=320, hq=200, fmt=0, k2=0.1, tilt=0.2)[0]
    ctx.set_hoist_chunk(8)
    try:
        got = common.detect_device(ctx, frames, common.to_vp(p))
    finally:
        ctx.set_hoist_chunk(0)
    assert got["plan"]["group"] == 8 and got["plan"]["reproject"] == 4 and got["plan"]["circ"] == 4, got["plan"]
    for i, fr in enumerate(frames):
        common.assert_frame_equal(got, i, port.detect(fr, p))


def test_tile_stats_of_the_headline_camera(ctx):
    """The undistorted top-down camera of the benchmark: every tile is staged, with 16-byte vectors (2448 % 16 == 0)."""
    p, _ = full_size_case(2448, 2048, n_frames=1)
    st = ctx.tile_stats(common.to_vp(p))
    assert st["tiles"] == 20 * 64 and st["staged"] == st["tiles"] and st["vectorised"] == st["tiles"], st


def test_misaligned_device_pointers_are_refused_not_faulted(ctx):
    """include/vp_b200.h states the alignment contract of the device-pointer API: a pointer that breaks it is refused with
    VP_ERR_INVALID before any kernel runs, and the context stays usable."""
    from vpb200 import lib
    p, raw, _ = common.make_case(wq=96, hq=64)
    vp = common.to_vp(p)
    nf = vp.wf * vp.hf
    b = dict(raw=ctx.buffer(raw.size + 64, None), flat=ctx.buffer(nf * 4 + 64), grad=ctx.buffer(nf * 4 + 64), circ=ctx.buffer(nf * 4 + 64),
             m=ctx.buffer(vp.max_blobs * 22), c=ctx.buffer(12))
    args = [b["raw"].device_ptr, 1, vp, b["flat"].device_ptr, b["grad"].device_ptr, b["circ"].device_ptr, b["m"].device_ptr, b["c"].device_ptr]
    for k in (0, 3, 4, 5):
        bad = list(args)
        bad[k] += 4
        with pytest.raises(lib.VpError) as e:
            ctx.detect_batch_device(*bad)
        assert e.value.code == 1 and "aligned" in str(e.value)
    ctx.to_device(b["raw"].device_ptr, raw)
    ctx.detect_batch_device(*args)  # the same context still works
    ctx.sync()
    for x in b.values():
        x.release()
