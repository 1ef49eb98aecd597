"""CPU: host-side geometry (the scalars the kernels are launched with) and the synthetic renderer."""
import math

import numpy as np
import pytest

from vpb200 import geometry as G, synth as S


def headline():
    cam = G.default_camera(1224, 1024)
    p = G.Perspective(cam)
    p.geometry_check(1224, 1024, 180.0)
    return cam, p


def test_headline_launch_parameters():
    """scale ~3.94 mm/px, flat == quad size, offset 2, r 6, radius 5."""
    cam, p = headline()
    lp = G.launch_params(p, 0, 1224, 1024)
    assert abs(p.field_scale - (5000 - 180) / 1224) < 1e-2
    assert (lp.wf, lp.hf) == (1224, 1024) and lp.wf % 2 == 0 and lp.hf % 2 == 0   # Perspective.cpp:118-122
    assert lp.grad_offset == int(math.ceil(25.0 / p.field_scale)) // 3 == 2          # Resources.cpp:160
    assert lp.circle_radius == int(math.ceil(20.0 / p.field_scale)) == 6             # Resources.cpp:163
    assert lp.blob_radius == int(math.floor(20.0 / p.field_scale)) == 5              # main.cpp:289
    assert lp.circ_threshold == 15.0 and lp.min_score == 0.0 and lp.max_blobs == 2000
    assert p.min_blob_radius == 20.0 and p.max_blob_radius == 25.0                   # Perspective.cpp:69-70


def test_cl_camera_model_layout():
    cam, _ = headline()
    b = G.pack_cl_camera_model(cam)
    assert len(b) == 72                                                               # Perspective.h:22-29
    assert np.frombuffer(b[0:8], "<i4").tolist() == [1224, 1024]
    assert np.frombuffer(b[8:12], "<f4")[0] == np.float32(1224.0)                    # focal length, not its reciprocal
    r = np.frombuffer(b[24:60], "<f4").reshape(3, 3)
    np.testing.assert_allclose(r @ r.T, np.eye(3), atol=1e-6)
    assert np.frombuffer(b[60:72], "<f4").tolist() == [0.0, 0.0, 5000.0]


def test_field2image_inverts_image2field():
    cam = G.CameraModel(size=(640, 480), focal_length=700.0, principal_point=(322.0, 238.0), distortion_k2=0.11, pos=(100.0, -50.0, 4000.0),
                        quat_wxyz=(0.05, -0.99, 0.02, 0.1))
    px = np.array([[10.0, 20.0], [320.0, 240.0], [600.0, 400.0]], np.float32)
    back = cam.field2image(cam.image2field(px, 150.0))
    np.testing.assert_allclose(back, px, atol=2e-2)


def test_flat_field_roundtrip():
    _, p = headline()
    pos = np.array([[0.0, 0.0], [100.5, 200.25]], np.float32)
    np.testing.assert_allclose(p.field2flat(p.flat2field(pos)), pos, atol=1e-3)


def test_renderer_is_deterministic_and_places_blobs():
    cam = G.default_camera(160, 120, height=180.0 + 4.0 * 160)
    p = G.Perspective(cam)
    p.geometry_check(160, 120, 180.0)
    sc = S.random_scene(p.visible_field_extent, 1, 1, seed=3)
    a = S.render_raw(sc, cam, 320, 240, S.FMT_RGGB, seed=5)
    b = S.render_raw(sc, cam, 320, 240, S.FMT_RGGB, seed=5)
    assert a.dtype == np.uint8 and a.shape == (240, 320) and np.array_equal(a, b)
    assert not np.array_equal(a, S.render_raw(sc, cam, 320, 240, S.FMT_RGGB, seed=6))
    gt = sc.ground_truth()                                                             # GroundTruth.cpp:23-78 schema
    assert len(gt["balls"]) == 1 and len(gt["robots_yellow"]) + len(gt["robots_blue"]) == 1
    assert len(sc.blobs()) == 6                                                        # 5 pattern blobs + 1 ball
    assert S.mosaic(np.zeros((4, 4, 3), np.uint8), S.FMT_BGR).shape == (4, 4, 3)


# ---- the C++ host derivation behind the C ABI (vision-processor_b200/host/geometry.cpp) against the numpy restatement ----
def _calib_of(cam: G.CameraModel):
    """SSL_GeometryCameraCalibration of a CameraModel, as CameraModel::getProto writes it (src/CameraModel.cpp:90-112)."""
    from vpb200 import lib
    w, x, y, z = cam.quat_wxyz
    t = cam.f2i() @ (-np.asarray(cam.pos, np.float32))
    return lib.CameraCalib(cam.size[0], cam.size[1], cam.focal_length, cam.principal_point[0], cam.principal_point[1], cam.distortion_k2,
                           x, y, z, w, float(t[0]), float(t[1]), float(t[2]))


def _field_c(f: G.FieldSize):
    from vpb200 import lib
    return lib.FieldSizeC(f.field_length, f.field_width, f.boundary_width, -1.0 if f.boundary_width_goal_line is None else f.boundary_width_goal_line,
                          f.ball_radius)


CAMERAS = [
    dict(cam=lambda: G.default_camera(1224, 1024), size=(1224, 1024)),
    dict(cam=lambda: G.default_camera(612, 512, k2=0.12), size=(612, 512)),
    dict(cam=lambda: G.CameraModel(size=(640, 480), focal_length=700.0, principal_point=(322.0, 238.0), distortion_k2=0.11, pos=(1500.0, -800.0, 4000.0),
                                   quat_wxyz=(0.05, -0.99, 0.02, 0.1)), size=(320, 240)),  # ensureSize halves it
]


@pytest.mark.parametrize("case", CAMERAS)
def test_host_geometry_matches_the_numpy_restatement(case):
    from vpb200 import lib
    cam = case["cam"]()
    ref = G.Perspective(case["cam"]())
    ref.geometry_check(*case["size"], 180.0, sequential_fp32=True)                     # the reference's own summation order
    hp = lib.HostPerspective(_calib_of(cam), _field_c(ref.field))
    hp.geometry_check(*case["size"], 180.0)
    assert hp.sees_field
    # the camera position survives the round trip through (q, t); the 72 kernel-argument bytes agree to rounding
    want = np.frombuffer(G.pack_cl_camera_model(ref.model), np.uint8)
    got = np.frombuffer(bytes(hp.model), np.uint8)
    assert got[:8].tolist() == want[:8].tolist()                                       # shape
    np.testing.assert_allclose(got[8:].view("<f4"), want[8:].view("<f4"), rtol=2e-6, atol=2e-3)
    # both accumulate the field scale sequentially in fp32 (Perspective.cpp:78-91); the terms differ by an ulp here and there
    assert abs(hp.field_scale - ref.field_scale) < 2e-5 * ref.field_scale
    np.testing.assert_allclose(hp.visible_field_extent, ref.visible_field_extent, rtol=1e-6, atol=2e-2)
    assert hp.reprojected_field_size == ref.reprojected_field_size
    assert (hp.min_blob_radius, hp.max_blob_radius) == (ref.min_blob_radius, ref.max_blob_radius)
    lp = G.launch_params(ref, 0, *case["size"])
    p = hp.params(0, *case["size"])
    assert (p.wf, p.hf, p.grad_offset, p.circle_radius, p.blob_radius, p.max_blobs) == (lp.wf, lp.hf, lp.grad_offset, lp.circle_radius, lp.blob_radius,
                                                                                      lp.max_blobs)
    assert p.min_score == 0.0 and p.circ_threshold == 15.0 and p.wf % 2 == 0 and p.hf % 2 == 0
    # point transforms
    px = np.array([[10.0, 20.0], [case["size"][0] / 2, case["size"][1] / 2], [case["size"][0] - 3.0, case["size"][1] - 7.0]], np.float32)
    fld = hp.image2field(px, 150.0)
    np.testing.assert_allclose(fld, ref.model.image2field(px, 150.0), rtol=1e-5, atol=5e-2)
    np.testing.assert_allclose(hp.field2image(fld), px, atol=3e-2)                     # 10 iterations invert the distortion
    np.testing.assert_allclose(hp.field2flat(hp.flat2field(px)), px, atol=1e-3)
    np.testing.assert_allclose(hp.flat2field(px), ref.flat2field(px), rtol=1e-5, atol=5e-2)


def test_host_geometry_of_a_camera_that_does_not_see_the_field():
    from vpb200 import lib
    cam = G.default_camera(64, 48)
    cam.pos = (50000.0, 0.0, 5000.0)                                                   # far outside the field
    hp = lib.HostPerspective(_calib_of(cam), _field_c(G.FieldSize()))
    hp.geometry_check(64, 48, 180.0)
    assert not hp.sees_field and math.isnan(hp.field_scale)                            # 0/0 like the reference
    with pytest.raises(lib.VpError):
        hp.params(0, 64, 48)


def test_host_geometry_rejects_bad_arguments():
    from vpb200 import lib
    bad = lib.CameraCalib(0, 0, -1.0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0)
    with pytest.raises(lib.VpError):
        lib.HostPerspective(bad, _field_c(G.FieldSize()))


def test_reference_summation_order_inflates_the_field_scale_at_full_size():
    """Perspective.cpp:78-91 adds 2.5 M distances of ~3.94 mm into one fp32: past 2^23 every `dx + dy` of 7.88 is absorbed
    as 8.  The C++ derivation follows the reference; the benchmark configuration uses the true mean."""
    from vpb200 import lib
    cam = G.default_camera(1224, 1024)
    true_mean = G.Perspective(G.default_camera(1224, 1024))
    true_mean.geometry_check(1224, 1024, 180.0)
    hp = lib.HostPerspective(_calib_of(cam), _field_c(G.FieldSize()))
    hp.geometry_check(1224, 1024, 180.0)
    assert abs(true_mean.field_scale - 4820 / 1224) < 1e-3 and true_mean.reprojected_field_size == (1224, 1024)
    assert 1.010 < hp.field_scale / true_mean.field_scale < 1.015
    assert hp.reprojected_field_size == (1208, 1010)                                   # instead of 1224 x 1024
