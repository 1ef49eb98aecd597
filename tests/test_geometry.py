"""CPU: host-side geometry (the scalars the kernels are launched with) and the synthetic renderer."""
import math

import numpy as np

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
