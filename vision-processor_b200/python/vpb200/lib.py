"""ctypes binding of libvp_b200.so (include/vp_b200.h) -- the product path.

Mirrors the reference's operator interface for the detection path: the names and argument order of
``Resources::raw2quad / rgba2blobCenter / quad2rgba / streamQuad / streamImage`` (src/Resources.cpp:138-186)
and the ``blobList`` launch of src/main.cpp:283-317.  Everything here goes through the C ABI; there is no
CPU implementation behind it and importing this module never touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Sequence

import numpy as np

from .geometry import LaunchParams

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(HERE, "..", "..", "lib", "libvp_b200.so"))
HEADER_PATH = os.path.normpath(os.path.join(HERE, "..", "..", "..", "include", "vp_b200.h"))

FMT_RGGB8, FMT_GRBG8, FMT_BGR8, FMT_RGBA8, FMT_U8, FMT_F32, FMT_NV12 = range(7)
SAMPLE_BILINEAR_RTE, SAMPLE_BILINEAR_TRUNC, SAMPLE_NEAREST = 0, 1, 2
MAP_READ, MAP_WRITE, MAP_READWRITE = 1, 2, 3

_NP_OF_FMT = {FMT_RGBA8: (np.uint8, 4), FMT_U8: (np.uint8, 1), FMT_F32: (np.float32, 1)}

# struct Match, src/blobs/match.h:22-30 (vp_field_match): what main.cpp:297-312 builds from every CLMatch
FIELD_MATCH_DTYPE = np.dtype([("pos", "<f4", 2), ("color", "<i4", 3), ("center", "<i4", 3), ("circ", "<f4"), ("score", "<f4")])
assert FIELD_MATCH_DTYPE.itemsize == 40

# CLMatch, src/main.cpp:33-41: 22 packed bytes, floats at unaligned offsets 14 and 18
MATCH_DTYPE = np.dtype({
    "names": ["x", "y", "color", "center", "circ", "score"],
    "formats": ["<f4", "<f4", ("u1", 3), ("u1", 3), "<f4", "<f4"],
    "offsets": [0, 4, 8, 11, 14, 18],
    "itemsize": 22,
})


class VpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vp_b200 error {code}: {msg}")
        self.code = code


class CameraModel(C.Structure):
    """vp_camera_model == CLCameraModel (src/Perspective.h:22-29), 72 packed bytes."""
    _pack_ = 1
    _fields_ = [("shape", C.c_int32 * 2), ("f", C.c_float), ("p", C.c_float * 2), ("d", C.c_float),
                ("r", C.c_float * 9), ("c", C.c_float * 3)]


class Params(C.Structure):
    """vp_params."""
    _fields_ = [
        ("fmt", C.c_int32), ("wq", C.c_int32), ("hq", C.c_int32), ("wf", C.c_int32), ("hf", C.c_int32),
        ("model", CameraModel),
        ("max_robot_height", C.c_float), ("field_scale", C.c_float), ("off_x", C.c_float), ("off_y", C.c_float),
        ("grad_offset", C.c_int32), ("circle_radius", C.c_int32),
        ("circ_threshold", C.c_float), ("min_score", C.c_float),
        ("blob_radius", C.c_int32), ("max_blobs", C.c_int32), ("sample_mode", C.c_int32),
    ]

    def raw_frame_bytes(self) -> int:
        return self.wq * self.hq * (3 if self.fmt == FMT_BGR8 else 4)


assert C.sizeof(CameraModel) == 72


class CameraCalib(C.Structure):
    """vp_camera_calib: the SSL_GeometryCameraCalibration fields CameraModel reads (src/CameraModel.cpp:80-88)."""
    _fields_ = [("pixel_image_width", C.c_int32), ("pixel_image_height", C.c_int32), ("focal_length", C.c_float),
                ("principal_point_x", C.c_float), ("principal_point_y", C.c_float), ("distortion", C.c_float),
                ("q0", C.c_float), ("q1", C.c_float), ("q2", C.c_float), ("q3", C.c_float),
                ("tx", C.c_float), ("ty", C.c_float), ("tz", C.c_float)]


class FieldSizeC(C.Structure):
    """vp_field_size."""
    _fields_ = [("field_length", C.c_float), ("field_width", C.c_float), ("boundary_width", C.c_float),
                ("boundary_width_goal_line", C.c_float), ("ball_radius", C.c_float)]


class Nv12Surface(C.Structure):
    """vp_nv12_surface: the layout an encoder session consumes (src/rtpstreamer.cpp:120-121)."""
    _fields_ = [("y", C.c_void_p), ("uv", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("pitch_y", C.c_int32), ("pitch_uv", C.c_int32),
                ("bytes_used", C.c_size_t), ("aligned16", C.c_int32)]


class Geometry(C.Structure):
    """vp_geometry: what Perspective::geometryCheck leaves behind (src/Perspective.h:32-58)."""
    _fields_ = [("model", CameraModel), ("field_scale", C.c_float), ("visible_field_extent", C.c_float * 4),
                ("reprojected_field_size", C.c_int32 * 2), ("min_blob_radius", C.c_float), ("max_blob_radius", C.c_float),
                ("min_field_scale", C.c_float), ("max_field_scale", C.c_float)]


def camera_model_from_bytes(b: bytes) -> CameraModel:
    assert len(b) == 72
    m = CameraModel()
    C.memmove(C.byref(m), b, 72)
    return m


def params_from_launch(lp: LaunchParams) -> Params:
    """geometry.LaunchParams -> vp_params (the scalars of Resources.cpp:159-163 and main.cpp:289)."""
    p = Params()
    p.fmt, p.wq, p.hq, p.wf, p.hf = lp.fmt, lp.wq, lp.hq, lp.wf, lp.hf
    p.model = camera_model_from_bytes(lp.model_bytes)
    p.max_robot_height, p.field_scale, p.off_x, p.off_y = lp.max_robot_height, lp.field_scale, lp.off_x, lp.off_y
    p.grad_offset, p.circle_radius = lp.grad_offset, lp.circle_radius
    p.circ_threshold, p.min_score = lp.circ_threshold, lp.min_score
    p.blob_radius, p.max_blobs, p.sample_mode = lp.blob_radius, lp.max_blobs, lp.sample_mode
    return p


def declared_symbols() -> list:
    """Every function include/vp_b200.h declares (VP_API lines)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"VP_API\s+[\w\s\*]+?\b(vp_\w+)\s*\(", text)))


_lib = None


def load() -> C.CDLL:
    """Load libvp_b200.so.  Fails loudly when the CUDA library has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    global LIB_PATH
    if os.environ.get("VPB200_LIB"):  # A/B builds of the same library (tools); never a fallback
        LIB_PATH = os.environ["VPB200_LIB"]
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(make -C vision-processor_b200/csrc).  vpb200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sigs = {
        "vp_version": (C.c_char_p, []),
        "vp_format_pixel_size": (C.c_int, [C.c_int]),
        "vp_camera_model_from_calib": (C.c_int, [C.POINTER(CameraCalib), C.POINTER(CameraModel)]),
        "vp_camera_model_ensure_size": (C.c_int, [C.POINTER(CameraModel), C.c_int, C.c_int]),
        "vp_field2image": (C.c_int, [C.POINTER(CameraModel), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "vp_image2field": (C.c_int, [C.POINTER(CameraModel), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]),
        "vp_geometry_check": (C.c_int, [C.POINTER(CameraModel), C.POINTER(FieldSizeC), C.c_int, C.c_int, C.c_double, C.c_float, C.c_float,
                                        C.POINTER(Geometry)]),
        "vp_flat2field": (C.c_int, [C.POINTER(Geometry), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "vp_field2flat": (C.c_int, [C.POINTER(Geometry), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "vp_geometry_params": (C.c_int, [C.POINTER(Geometry), C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_int, C.c_int, C.POINTER(Params)]),
        "vp_device_count": (C.c_int, []),
        "vp_last_error": (C.c_char_p, [vp]),
        "vp_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "vp_ctx_destroy": (None, [vp]),
        "vp_ctx_sync": (C.c_int, [vp]),
        "vp_ctx_stream": (vp, [vp]),
        "vp_ctx_set_group": (C.c_int, [vp, C.c_int]),
        "vp_ctx_set_lanes": (C.c_int, [vp, C.c_int]),
        "vp_ctx_set_hoist_chunk": (C.c_int, [vp, C.c_int]),
        "vp_ctx_set_strips": (C.c_int, [vp, C.c_int]),
        "vp_ctx_set_latency_graph": (C.c_int, [vp, C.c_int]),
        "vp_latency_graph_replays": (C.c_uint64, [vp]),
        "vp_ctx_set_staged_reproject": (C.c_int, [vp, C.c_int]),
        "vp_ctx_set_fused_gradcirc": (C.c_int, [vp, C.c_int]),
        "vp_launch_count": (C.c_uint64, [vp]),
        "vp_detect_last_plan": (C.c_int, [vp, C.POINTER(C.c_int32)]),
        "vp_tile_stats": (C.c_int, [vp, C.POINTER(Params), C.POINTER(C.c_int32)]),
        "vp_profiling_enable": (C.c_int, [vp, C.c_int]),
        "vp_profiling_count": (C.c_int, [vp]),
        "vp_profiling_get": (C.c_int, [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]),
        "vp_profiling_clear": (C.c_int, [vp]),
        "vp_buf_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
        "vp_buf_alloc_copy": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(vp)]),
        "vp_buf_retain": (C.c_int, [vp]),
        "vp_buf_release": (C.c_int, [vp]),
        "vp_buf_map": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
        "vp_buf_unmap": (C.c_int, [vp]),
        "vp_buf_size": (C.c_size_t, [vp]),
        "vp_buf_device_ptr": (vp, [vp]),
        "vp_img_alloc": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
        "vp_img_retain": (C.c_int, [vp]),
        "vp_img_release": (C.c_int, [vp]),
        "vp_img_map": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "vp_img_unmap": (C.c_int, [vp]),
        "vp_img_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "vp_img_device_ptr": (vp, [vp]),
        "vp_raw2quad": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
        "vp_resampling": (C.c_int, [vp, C.POINTER(vp), C.c_int, vp, C.POINTER(CameraModel), C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]),
        "vp_gradient_dot": (C.c_int, [vp, vp, vp, C.c_int]),
        "vp_sat_horizontal": (C.c_int, [vp, vp, vp]),
        "vp_sat_vertical": (C.c_int, [vp, vp, vp]),
        "vp_circle": (C.c_int, [vp, vp, vp, C.c_int]),
        "vp_blob_list": (C.c_int, [vp, vp, vp, vp, vp, C.c_float, C.c_float, C.c_int, C.c_int]),
        "vp_rgba2nv12": (C.c_int, [vp, vp, vp]),
        "vp_f2nv12": (C.c_int, [vp, vp, vp]),
        "vp_quad2nv12": (C.c_int, [vp, C.POINTER(vp), C.c_int, vp, C.c_int]),
        "vp_quad2rgba": (C.c_int, [vp, C.POINTER(vp), C.c_int, vp, C.c_int]),
        "vp_circularize": (C.c_int, [vp, vp, vp, C.c_int, C.c_int]),
        "vp_blob_score": (C.c_int, [vp, vp, vp, vp, C.c_float, C.c_int]),
        "vp_detect_batch_device": (C.c_int, [vp, vp, C.c_int, C.POINTER(Params), vp, vp, vp, vp, vp]),
        "vp_detect_host": (C.c_int, [vp, vp, C.c_int, C.POINTER(Params), vp, vp]),
        "vp_detect_images": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "vp_detect_sat_fallbacks": (C.c_int, [vp, C.POINTER(C.c_int)]),
        "vp_raw2nv12_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
        "vp_raw2rgba_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
        "vp_rgba2nv12_device": (C.c_int, [vp, vp, C.c_int, C.c_int, vp]),
        "vp_f2nv12_device": (C.c_int, [vp, vp, C.c_int, C.c_int, vp]),
        "vp_rgba2nv12_batch_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t]),
        "vp_f2nv12_batch_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t]),
        "vp_raw2nv12_batch_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, C.c_int]),
        "vp_blobs_to_field_device": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, vp, vp, vp]),
        "vp_nv12_surface_of": (C.c_int, [vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(Nv12Surface)]),
        "vp_copy_to_host": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "vp_copy_to_device": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "vp_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
        "vp_host_free": (C.c_int, [vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def bound_symbols() -> list:
    load()
    return sorted(n for n in declared_symbols() if hasattr(_lib, n))


def device_count() -> int:
    return int(load().vp_device_count())


def _ck_host(rc: int, what: str, allow=()) -> int:
    if rc != 0 and rc not in allow:
        raise VpError(rc, what)
    return rc


class HostPerspective:
    """The host-side derivation behind the C ABI (vision-processor_b200/host/geometry.cpp), with the member names of the
    reference's ``Perspective`` (src/Perspective.h:32-58): model, fieldScale, visibleFieldExtent, reprojectedFieldSize,
    minBlobRadius, maxBlobRadius; geometryCheck, flat2field, field2flat.  CPU only, no context needed."""

    def __init__(self, calib: CameraCalib, field: FieldSizeC, geometry_tolerance: float = 10.0):
        self.lib = load()
        self.model = CameraModel()
        _ck_host(self.lib.vp_camera_model_from_calib(C.byref(calib), C.byref(self.model)), "vp_camera_model_from_calib")
        self.field = field
        self.geometry_tolerance = geometry_tolerance
        self.geometry = Geometry()
        self.sees_field = False

    def geometry_check(self, width: int, height: int, max_bot_height: float, resampling_factor: float = 1.0) -> None:
        rc = _ck_host(self.lib.vp_geometry_check(C.byref(self.model), C.byref(self.field), width, height, max_bot_height, resampling_factor,
                                                 self.geometry_tolerance, C.byref(self.geometry)), "vp_geometry_check", allow=(4,))
        self.sees_field = rc == 0
        self.model = self.geometry.model

    field_scale = property(lambda self: float(self.geometry.field_scale))
    visible_field_extent = property(lambda self: tuple(self.geometry.visible_field_extent))
    reprojected_field_size = property(lambda self: tuple(self.geometry.reprojected_field_size))
    min_blob_radius = property(lambda self: float(self.geometry.min_blob_radius))
    max_blob_radius = property(lambda self: float(self.geometry.max_blob_radius))

    def _xy(self, fn, a, n_in: int, n_out: int, *extra) -> np.ndarray:
        a = np.asarray(a, np.float32).reshape(-1, n_in)
        out = np.empty((len(a), n_out), np.float32)
        for i in range(len(a)):
            src, dst = (C.c_float * n_in)(*a[i]), (C.c_float * n_out)()
            _ck_host(fn(*extra[:1], src, *extra[1:], dst), fn.__name__)
            out[i] = list(dst)
        return out

    def flat2field(self, pos) -> np.ndarray:
        return self._xy(self.lib.vp_flat2field, pos, 2, 2, C.byref(self.geometry))

    def field2flat(self, pos) -> np.ndarray:
        return self._xy(self.lib.vp_field2flat, pos, 2, 2, C.byref(self.geometry))

    def field2image(self, pos) -> np.ndarray:
        return self._xy(self.lib.vp_field2image, pos, 3, 2, C.byref(self.model))

    def image2field(self, pos, height: float) -> np.ndarray:
        return self._xy(self.lib.vp_image2field, pos, 2, 3, C.byref(self.model), C.c_float(height))

    def params(self, fmt: int, wq: int, hq: int, max_bot_height: float = 180.0, circ_threshold: float = 15.0, max_blobs: int = 2000,
               sample_mode: int = 0) -> Params:
        p = Params()
        _ck_host(self.lib.vp_geometry_params(C.byref(self.geometry), fmt, wq, hq, max_bot_height, circ_threshold, max_blobs, sample_mode, C.byref(p)),
                 "vp_geometry_params")
        return p


def _np_ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class PinnedArray:
    """A numpy view of pinned host memory (vp_host_alloc): the camera drivers' user buffers."""

    def __init__(self, shape, dtype=np.uint8):
        self.lib = load()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = self.lib.vp_host_alloc(n, C.byref(p))
        if rc:
            raise VpError(rc, self.lib.vp_last_error(None).decode())
        self.ptr = p
        self.array = np.frombuffer((C.c_uint8 * max(n, 1)).from_address(p.value), dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.vp_host_free(self.ptr)
            self.ptr = None


class Image:
    """vp_img handle (class CLImage, src/opencl.h:195-212)."""

    def __init__(self, ctx: "Context", fmt: int, width: int, height: int):
        self.ctx, self.fmt, self.width, self.height = ctx, fmt, width, height
        h = C.c_void_p()
        ctx._ck(ctx.lib.vp_img_alloc(ctx.h, fmt, width, height, C.byref(h)))
        self.h = h

    @property
    def device_ptr(self) -> int:
        return int(self.ctx.lib.vp_img_device_ptr(self.h) or 0)

    def _shape(self):
        dt, ch = _NP_OF_FMT[self.fmt]
        return dt, ((self.height, self.width, ch) if ch > 1 else (self.height, self.width))

    def write(self, a: np.ndarray) -> "Image":
        dt, shape = self._shape()
        a = np.ascontiguousarray(a, dt).reshape(shape)
        p, pitch = C.c_void_p(), C.c_size_t()
        self.ctx._ck(self.ctx.lib.vp_img_map(self.h, MAP_WRITE, C.byref(p), C.byref(pitch)))
        if a.nbytes:
            C.memmove(p, a.ctypes.data, a.nbytes)
        self.ctx._ck(self.ctx.lib.vp_img_unmap(self.h))
        return self

    def read(self) -> np.ndarray:
        dt, shape = self._shape()
        out = np.empty(shape, dt)
        p, pitch = C.c_void_p(), C.c_size_t()
        self.ctx._ck(self.ctx.lib.vp_img_map(self.h, MAP_READ, C.byref(p), C.byref(pitch)))
        assert pitch.value == self.width * out.itemsize * (shape[2] if len(shape) == 3 else 1)
        if out.nbytes:
            C.memmove(out.ctypes.data, p, out.nbytes)
        self.ctx._ck(self.ctx.lib.vp_img_unmap(self.h))
        return out

    def release(self):
        if self.h:
            self.ctx.lib.vp_img_release(self.h)
            self.h = None


class Buffer:
    """vp_buf handle (class CLArray, src/opencl.h:154-165)."""

    def __init__(self, ctx: "Context", nbytes: int, data: np.ndarray | None = None):
        self.ctx, self.nbytes = ctx, int(nbytes)
        h = C.c_void_p()
        if data is not None:
            data = np.ascontiguousarray(data)
            assert data.nbytes == self.nbytes
            ctx._ck(ctx.lib.vp_buf_alloc_copy(ctx.h, _np_ptr(data), self.nbytes, C.byref(h)))
        else:
            ctx._ck(ctx.lib.vp_buf_alloc(ctx.h, self.nbytes, C.byref(h)))
        self.h = h

    @property
    def device_ptr(self) -> int:
        return int(self.ctx.lib.vp_buf_device_ptr(self.h) or 0)

    def write(self, a: np.ndarray) -> "Buffer":
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        p = C.c_void_p()
        self.ctx._ck(self.ctx.lib.vp_buf_map(self.h, MAP_READWRITE if a.nbytes < self.nbytes else MAP_WRITE, C.byref(p)))
        if a.nbytes:
            C.memmove(p, a.ctypes.data, a.nbytes)
        self.ctx._ck(self.ctx.lib.vp_buf_unmap(self.h))
        return self

    def read(self, dtype=np.uint8, count: int | None = None) -> np.ndarray:
        dt = np.dtype(dtype)
        n = self.nbytes // dt.itemsize if count is None else count
        out = np.empty(n, dt)
        p = C.c_void_p()
        self.ctx._ck(self.ctx.lib.vp_buf_map(self.h, MAP_READ, C.byref(p)))
        if out.nbytes:
            C.memmove(out.ctypes.data, p, out.nbytes)
        self.ctx._ck(self.ctx.lib.vp_buf_unmap(self.h))
        return out

    def release(self):
        if self.h:
            self.ctx.lib.vp_buf_release(self.h)
            self.h = None


class Context:
    """vp_ctx (class OpenCL, src/opencl.h:69-112) plus numpy-in/numpy-out wrappers of every stage."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.vp_ctx_create(device, C.byref(h))
        if rc:
            raise VpError(rc, self.lib.vp_last_error(None).decode())
        self.h = h
        self.device = device

    def _ck(self, rc: int):
        if rc:
            raise VpError(rc, self.lib.vp_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.vp_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        self._ck(self.lib.vp_ctx_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.vp_ctx_stream(self.h) or 0)

    def launch_count(self) -> int:
        return int(self.lib.vp_launch_count(self.h))

    def set_group(self, n: int):
        self._ck(self.lib.vp_ctx_set_group(self.h, n))

    def set_staged_reproject(self, on):
        """0/False = direct gather, 1/True or 2 = shared-memory staged with the frame-invariant part hoisted (default)."""
        self._ck(self.lib.vp_ctx_set_staged_reproject(self.h, int(on)))

    def set_fused_gradcirc(self, on):
        """One gradient + circularity kernel vs gradient + row sums followed by the streaming circularity kernel: 0/False never,
        1/True for calls of more than two frames (default), 2 also for the one- and two-frame calls of the latency path."""
        self._ck(self.lib.vp_ctx_set_fused_gradcirc(self.h, int(on)))

    def set_hoist_chunk(self, n: int):
        self._ck(self.lib.vp_ctx_set_hoist_chunk(self.h, n))

    def last_plan(self) -> dict:
        """What the most recent fused call launched (vp_detect_last_plan): which specialised kernels a test really exercised."""
        a = (C.c_int32 * 8)()
        self._ck(self.lib.vp_detect_last_plan(self.h, a))
        return dict(reproject=a[0], chunk=a[1], group=a[2], lanes=a[3], circ=a[4], seg_rows=a[5], tma=a[6])

    def tile_stats(self, p) -> dict:
        """How the geometry maps onto the staged reprojection (vp_tile_stats): tiles per frame, tiles that fit the staged planes."""
        a = (C.c_int32 * 4)()
        self._ck(self.lib.vp_tile_stats(self.h, C.byref(p), a))
        return dict(tiles=a[0], staged=a[1], vectorised=a[2], max_rows=a[3])

    def set_strips(self, n: int):
        """Chunks the upload of a lone frame is cut into on the latency path of detect_host (1 = no overlap)."""
        self._ck(self.lib.vp_ctx_set_strips(self.h, n))

    def set_latency_graph(self, on: bool):
        """Replay the one-frame sequence of detect_host as a CUDA graph (default on)."""
        self._ck(self.lib.vp_ctx_set_latency_graph(self.h, int(on)))

    def latency_graph_replays(self) -> int:
        return int(self.lib.vp_latency_graph_replays(self.h))

    def set_lanes(self, n: int):
        self._ck(self.lib.vp_ctx_set_lanes(self.h, n))

    # ---- profiling (OpenCL::printRuntimes, opencl.cpp:94-101) -----------------------------------
    def profiling(self, on: bool):
        self._ck(self.lib.vp_profiling_enable(self.h, int(on)))

    def runtimes(self, clear: bool = True) -> list:
        out = []
        name, ms = C.c_char_p(), C.c_float()
        for i in range(self.lib.vp_profiling_count(self.h)):
            self._ck(self.lib.vp_profiling_get(self.h, i, C.byref(name), C.byref(ms)))
            out.append((name.value.decode(), float(ms.value)))
        if clear:
            self._ck(self.lib.vp_profiling_clear(self.h))
        return out

    # ---- helpers --------------------------------------------------------------------------------
    def image(self, fmt, w, h) -> Image:
        return Image(self, fmt, w, h)

    def image_from(self, a: np.ndarray) -> Image:
        if a.dtype == np.float32:
            fmt = FMT_F32
        elif a.ndim == 3 and a.shape[2] == 4:
            fmt = FMT_RGBA8
        else:
            fmt = FMT_U8
        return Image(self, fmt, a.shape[1], a.shape[0]).write(a)

    def buffer(self, nbytes, data=None) -> Buffer:
        return Buffer(self, nbytes, data)

    def to_host(self, dev_ptr: int, nbytes: int, dtype=np.uint8) -> np.ndarray:
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype)
        self._ck(self.lib.vp_copy_to_host(self.h, _np_ptr(out), C.c_void_p(dev_ptr), out.nbytes))
        return out

    def to_device(self, dev_ptr: int, a: np.ndarray):
        a = np.ascontiguousarray(a)
        self._ck(self.lib.vp_copy_to_device(self.h, C.c_void_p(dev_ptr), _np_ptr(a), a.nbytes))

    @staticmethod
    def _handles(imgs: Sequence[Image]):
        arr = (C.c_void_p * 4)()
        for i in range(4):
            arr[i] = imgs[i].h if i < len(imgs) and imgs[i] is not None else None
        return arr

    # ---- stages, numpy in / numpy out -----------------------------------------------------------
    def raw2quad(self, raw: np.ndarray, fmt: int, wq: int, hq: int, stale: int = 0):
        """Resources::raw2quad (Resources.cpp:138-143).  Returns the four U8 planes (BGR leaves plane 3 = `stale`)."""
        buf = self.buffer(raw.nbytes, raw)
        ch = [self.image_from(np.full((hq, wq), stale, np.uint8)) for _ in range(4)]
        self._ck(self.lib.vp_raw2quad(self.h, buf.h, fmt, wq, hq, self._handles(ch)))
        out = [c.read() for c in ch]
        for c in ch:
            c.release()
        buf.release()
        return out

    def resampling(self, ch, fmt, wf, hf, model: CameraModel, height, scale, offx, offy, mode=0):
        imgs = [self.image_from(np.ascontiguousarray(c)) for c in ch]
        flat = self.image(FMT_RGBA8, wf, hf)
        self._ck(self.lib.vp_resampling(self.h, self._handles(imgs), fmt, flat.h, C.byref(model), height, scale, offx, offy, mode))
        out = flat.read()
        for i in imgs + [flat]:
            i.release()
        return out

    def _unary(self, fn, a: np.ndarray, out_fmt: int, *extra):
        i = self.image_from(a)
        o = self.image(out_fmt, i.width, i.height)
        self._ck(fn(self.h, i.h, o.h, *extra))
        out = o.read()
        i.release()
        o.release()
        return out

    def gradient_dot(self, rgba, offset):
        return self._unary(self.lib.vp_gradient_dot, np.ascontiguousarray(rgba, np.uint8), FMT_F32, int(offset))

    def sat_horizontal(self, a):
        return self._unary(self.lib.vp_sat_horizontal, np.ascontiguousarray(a, np.float32), FMT_F32)

    def sat_vertical(self, a):
        return self._unary(self.lib.vp_sat_vertical, np.ascontiguousarray(a, np.float32), FMT_F32)

    def circle(self, sat, r):
        return self._unary(self.lib.vp_circle, np.ascontiguousarray(sat, np.float32), FMT_F32, int(r))

    def circularize(self, a, minr, maxr):
        return self._unary(self.lib.vp_circularize, np.ascontiguousarray(a, np.float32), FMT_F32, int(minr), int(maxr))

    def blob_list(self, rgba, circ, thr, min_score, radius, max_matches, counter0=(0, 0, 0)):
        """The blobList launch + readback of main.cpp:283-317."""
        i = self.image_from(np.ascontiguousarray(rgba, np.uint8))
        c = self.image_from(np.ascontiguousarray(circ, np.float32))
        m = self.buffer(max(max_matches, 1) * 22)
        cnt = self.buffer(12, np.asarray(counter0, np.int32))
        self._ck(self.lib.vp_blob_list(self.h, i.h, c.h, m.h, cnt.h, thr, min_score, int(radius), int(max_matches)))
        counter = cnt.read(np.int32)
        first = int(counter0[0])
        n = max(0, min(int(counter[0]), max_matches) - first)
        raw = m.read(np.uint8)
        matches = raw[22 * first: 22 * (first + n)].view(MATCH_DTYPE).copy()
        for x in (i, c, m, cnt):
            x.release()
        return matches, counter

    def blob_score(self, rgba, circ, thr, radius):
        i = self.image_from(np.ascontiguousarray(rgba, np.uint8))
        c = self.image_from(np.ascontiguousarray(circ, np.float32))
        o = self.image(FMT_F32, i.width, i.height)
        self._ck(self.lib.vp_blob_score(self.h, i.h, c.h, o.h, thr, int(radius)))
        out = o.read()
        for x in (i, c, o):
            x.release()
        return out

    def _to_nv12(self, fn, a):
        i = self.image_from(a)
        buf = self.buffer(2 * i.width * i.height, np.zeros(2 * i.width * i.height, np.uint8))
        self._ck(fn(self.h, i.h, buf.h))
        out = buf.read()
        i.release()
        buf.release()
        return out

    def rgba2nv12(self, rgba):
        """Resources::streamImage for RGBA8 (Resources.cpp:172-186).  Returns the 2*w*h NV12 buffer (1.5*w*h used)."""
        return self._to_nv12(self.lib.vp_rgba2nv12, np.ascontiguousarray(rgba, np.uint8))

    def f2nv12(self, a):
        return self._to_nv12(self.lib.vp_f2nv12, np.ascontiguousarray(a, np.float32))

    def quad2nv12(self, ch, fmt, mode=0):
        """Resources::streamQuad (Resources.cpp:166-170)."""
        imgs = [self.image_from(np.ascontiguousarray(c)) for c in ch]
        hq, wq = ch[0].shape
        buf = self.buffer(2 * wq * hq, np.zeros(2 * wq * hq, np.uint8))
        self._ck(self.lib.vp_quad2nv12(self.h, self._handles(imgs), fmt, buf.h, mode))
        out = buf.read()
        for i in imgs:
            i.release()
        buf.release()
        return out

    def quad2rgba(self, ch, fmt, mode=0):
        """Resources::quad2rgba (Resources.cpp:145-149)."""
        imgs = [self.image_from(np.ascontiguousarray(c)) for c in ch]
        hq, wq = ch[0].shape
        o = self.image(FMT_RGBA8, wq, hq)
        self._ck(self.lib.vp_quad2rgba(self.h, self._handles(imgs), fmt, o.h, mode))
        out = o.read()
        for i in imgs + [o]:
            i.release()
        return out

    # ---- fused detection ------------------------------------------------------------------------
    def detect_batch_device(self, d_raw: int, n_frames: int, p: Params, d_flat: int, d_grad: int, d_circ: int, d_matches: int, d_counter: int):
        """Asynchronous; all arguments are device addresses (e.g. torch.Tensor.data_ptr())."""
        self._ck(self.lib.vp_detect_batch_device(self.h, C.c_void_p(d_raw), n_frames, C.byref(p), C.c_void_p(d_flat), C.c_void_p(d_grad),
                                                 C.c_void_p(d_circ), C.c_void_p(d_matches), C.c_void_p(d_counter)))

    def detect_host_into(self, h_raw_ptr: int, n_frames: int, p: Params, h_matches_ptr: int, h_counter_ptr: int):
        """Blocking; raw frames, blob lists and counters in (pinned) host memory."""
        self._ck(self.lib.vp_detect_host(self.h, C.c_void_p(h_raw_ptr), n_frames, C.byref(p), C.c_void_p(h_matches_ptr), C.c_void_p(h_counter_ptr)))

    def sat_fallbacks(self) -> int:
        n = C.c_int()
        self._ck(self.lib.vp_detect_sat_fallbacks(self.h, C.byref(n)))
        return int(n.value)

    def detect(self, raw: np.ndarray, p: Params, want_images: bool = True, pinned: bool = False) -> dict:
        """One or more frames through raw2quad + rgba2blobCenter + blobList (Resources.cpp:138-164, main.cpp:283-317).

        raw: (n, raw_bytes) or a single frame; pinned=True stages it in pinned host memory first (the latency path of a
        lone frame only overlaps the upload with the kernels for pinned frames).  Returns the blob lists (clamped to max_blobs like main.cpp:301),
        the counters, and -- for the last frame -- the `flat`, `gradDot`, `blobCenter` images."""
        rb = p.raw_frame_bytes()
        raw = np.ascontiguousarray(raw, np.uint8).reshape(-1, rb)
        n = raw.shape[0]
        matches = np.zeros((n, max(p.max_blobs, 1)), MATCH_DTYPE)
        counter = np.zeros((n, 3), np.int32)
        if pinned:  # the frames in pinned host memory, like a camera driver's buffers: upload in strips, graph replay
            ring = PinnedArray(raw.shape, np.uint8)
            try:
                ring.array[:] = raw
                self.detect_host_into(ring.ptr.value, n, p, matches.ctypes.data, counter.ctypes.data)
            finally:
                ring.free()
        else:
            self.detect_host_into(raw.ctypes.data, n, p, matches.ctypes.data, counter.ctypes.data)
        out = {
            "matches": [matches[i, : min(int(counter[i, 0]), p.max_blobs)].copy() for i in range(n)],
            "counter": counter,
            "sat_fallbacks": self.sat_fallbacks(),
        }
        if want_images:
            f, g, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
            self._ck(self.lib.vp_detect_images(self.h, C.byref(f), C.byref(g), C.byref(c)))
            nf = p.wf * p.hf
            out["flat"] = self.to_host(f.value, nf * 4).reshape(p.hf, p.wf, 4)
            out["grad"] = self.to_host(g.value, nf * 4, np.float32).reshape(p.hf, p.wf)
            out["circ"] = self.to_host(c.value, nf * 4, np.float32).reshape(p.hf, p.wf)
        return out

    def raw2nv12(self, raw: np.ndarray, fmt, wq, hq, mode=0):
        """streamQuad straight from the raw frame (no quad planes)."""
        src = self.buffer(raw.nbytes, raw)
        dst = self.buffer(2 * wq * hq, np.zeros(2 * wq * hq, np.uint8))
        self._ck(self.lib.vp_raw2nv12_device(self.h, C.c_void_p(src.device_ptr), fmt, wq, hq, C.c_void_p(dst.device_ptr), mode))
        out = dst.read()
        src.release()
        dst.release()
        return out

    def nv12_batch(self, kind: str, frames: np.ndarray, w: int, h: int, fmt: int = 0, mode: int = 0, stride: int | None = None) -> np.ndarray:
        """n frames -> n NV12 views in one launch.  kind: 'rgba' (n,h,w,4 u8), 'f32' (n,h,w f32) or 'raw' (n raw frames)."""
        frames = np.ascontiguousarray(frames)
        n = len(frames)
        stride = 2 * w * h if stride is None else stride
        src = self.buffer(max(frames.nbytes, 1), frames.reshape(-1).view(np.uint8))
        dst = self.buffer(max(n * stride, 1), np.zeros(max(n * stride, 1), np.uint8))
        if kind == "rgba":
            rc = self.lib.vp_rgba2nv12_batch_device(self.h, C.c_void_p(src.device_ptr), n, w, h, C.c_void_p(dst.device_ptr), stride)
        elif kind == "f32":
            rc = self.lib.vp_f2nv12_batch_device(self.h, C.c_void_p(src.device_ptr), n, w, h, C.c_void_p(dst.device_ptr), stride)
        else:
            rc = self.lib.vp_raw2nv12_batch_device(self.h, C.c_void_p(src.device_ptr), n, fmt, w, h, C.c_void_p(dst.device_ptr), stride, mode)
        try:
            self._ck(rc)
            return dst.read()[:n * stride].reshape(n, stride)
        finally:
            src.release()
            dst.release()

    def blobs_to_field(self, matches: np.ndarray, counters: np.ndarray, max_blobs: int, field_scale: float, off_x: float, off_y: float,
                       cell_mm: float, cells_x: int, cells_y: int):
        """main.cpp:297-325 on the device for a batch: matches (n, max_blobs) CLMatch records, counters (n, 3).
        Returns (records (n, max_blobs) FIELD_MATCH_DTYPE, order (n, max_blobs) i32, cell_start (n, cells+1) i32)."""
        n = len(counters)
        m = self.buffer(max(n * max_blobs * 22, 1), np.ascontiguousarray(matches).view(np.uint8).reshape(-1))
        c = self.buffer(n * 12, np.ascontiguousarray(counters, np.int32).view(np.uint8).reshape(-1))
        out = self.buffer(max(n * max_blobs * 40, 1))
        order = self.buffer(max(n * max_blobs * 4, 1))
        cs = self.buffer(n * (cells_x * cells_y + 1) * 4)
        try:
            self._ck(self.lib.vp_blobs_to_field_device(self.h, C.c_void_p(m.device_ptr), C.c_void_p(c.device_ptr), n, max_blobs, field_scale, off_x, off_y,
                                                       cell_mm, cells_x, cells_y, C.c_void_p(out.device_ptr), C.c_void_p(order.device_ptr),
                                                       C.c_void_p(cs.device_ptr)))
            rec = out.read()[:n * max_blobs * 40].view(FIELD_MATCH_DTYPE).reshape(n, max_blobs)
            return rec, order.read(np.int32)[:n * max_blobs].reshape(n, max_blobs), cs.read(np.int32).reshape(n, cells_x * cells_y + 1)
        finally:
            for b in (m, c, out, order, cs):
                b.release()

    def raw2rgba(self, raw: np.ndarray, fmt, wq, hq, mode=0):
        src = self.buffer(raw.nbytes, raw)
        dst = self.buffer(4 * wq * hq)
        self._ck(self.lib.vp_raw2rgba_device(self.h, C.c_void_p(src.device_ptr), fmt, wq, hq, C.c_void_p(dst.device_ptr), mode))
        out = dst.read().reshape(hq, wq, 4)
        src.release()
        dst.release()
        return out


def canonical(matches: np.ndarray) -> np.ndarray:
    """Canonical ordering of a blob list: byte-wise sort of the 22-byte records."""
    raw = np.ascontiguousarray(matches).view(np.uint8).reshape(-1, 22)
    return matches[np.lexsort(raw.T[::-1])]
