"""ctypes front-end for the CPU ORACLE (test infrastructure, NOT product code).

Two interchangeable back-ends with the same call surface:

* ``Oracle("port")``      -> oracle/libvp_oracle.so   (vp_oracle.c, the plain-C restatement)
* ``Oracle("reference")`` -> oracle/_ref/libvp_clref.so (the reference's own kernel/*.cl,
  compiled in place through clemu.h; only present when it was built in the container)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (vpb200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

FMT_RGGB, FMT_GRBG, FMT_BGR = 0, 1, 2
SAMPLE_BILINEAR_RTE, SAMPLE_BILINEAR_TRUNC, SAMPLE_NEAREST = 0, 1, 2

# blobList.cl:20-32 / main.cpp:33-41 -- 22 bytes, floats at unaligned offsets 14 and 18
MATCH_DTYPE = np.dtype(
    {
        "names": ["x", "y", "color", "center", "circ", "score"],
        "formats": ["<f4", "<f4", ("u1", 3), ("u1", 3), "<f4", "<f4"],
        "offsets": [0, 4, 8, 11, 14, 18],
        "itemsize": 22,
    }
)


class CameraModel(C.Structure):
    """resampling.cl:20-27 == Perspective.h:22-29 (packed, 72 bytes)."""

    _pack_ = 1
    _fields_ = [
        ("shape", C.c_int32 * 2),
        ("f", C.c_float),
        ("p", C.c_float * 2),
        ("d", C.c_float),
        ("r", C.c_float * 9),
        ("c", C.c_float * 3),
    ]


class Params(C.Structure):
    """vpo_params (vp_oracle.h)."""

    _fields_ = [
        ("fmt", C.c_int), ("wq", C.c_int), ("hq", C.c_int), ("wf", C.c_int), ("hf", C.c_int),
        ("model", CameraModel),
        ("max_robot_height", C.c_float), ("field_scale", C.c_float), ("off_x", C.c_float), ("off_y", C.c_float),
        ("grad_offset", C.c_int), ("circle_radius", C.c_int),
        ("circ_threshold", C.c_float), ("min_score", C.c_float),
        ("blob_radius", C.c_int), ("max_blobs", C.c_int), ("sample_mode", C.c_int),
    ]


assert C.sizeof(CameraModel) == 72


def build(verbose: bool = False) -> None:
    """Run oracle/Makefile (libvp_oracle.so always; _ref/ only if /root/reference exists)."""
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def have_reference() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libvp_clref.so"))


class Oracle:
    def __init__(self, kind: str = "port"):
        self.kind = kind
        if kind == "port":
            path, self.pfx = os.path.join(HERE, "libvp_oracle.so"), "vpo_"
            if not os.path.exists(path):
                build()
        elif kind == "reference":
            path, self.pfx = os.path.join(HERE, "_ref", "libvp_clref.so"), "clref_"
        else:
            raise ValueError(kind)
        self.lib = C.CDLL(path)
        self._f("detect").restype = C.c_float

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    def set_threads(self, n: int) -> None:
        self._f("set_threads")(C.c_int(int(n)))

    # ---- stages -------------------------------------------------------------------------
    def raw2quad(self, raw: np.ndarray, fmt: int, wq: int, hq: int, stale: int = 0):
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        ch = [np.full((hq, wq), stale, np.uint8) for _ in range(4)]
        self._f("raw2quad")(_u8(raw), fmt, wq, hq, *[_u8(c) for c in ch])
        return ch

    def resampling(self, ch, fmt, wf, hf, model: CameraModel, height, scale, offx, offy, mode=0):
        hq, wq = ch[0].shape
        flat = np.zeros((hf, wf, 4), np.uint8)
        self._f("resampling")(*[_u8(np.ascontiguousarray(c)) for c in ch], fmt, wq, hq, _u8(flat), wf, hf,
                              C.byref(model), C.c_float(height), C.c_float(scale), C.c_float(offx), C.c_float(offy), mode)
        return flat

    def gradient_dot(self, rgba, offset):
        h, w = rgba.shape[:2]
        out = np.zeros((h, w), np.float32)
        self._f("gradient_dot")(_u8(np.ascontiguousarray(rgba)), w, h, int(offset), _f32(out))
        return out

    def sat_horizontal(self, a):
        a = np.ascontiguousarray(a, np.float32)
        out = np.zeros_like(a)
        self._f("sat_horizontal")(_f32(a), a.shape[1], a.shape[0], _f32(out))
        return out

    def sat_vertical(self, a):
        a = np.ascontiguousarray(a, np.float32)
        out = np.zeros_like(a)
        self._f("sat_vertical")(_f32(a), a.shape[1], a.shape[0], _f32(out))
        return out

    def circle(self, sat, r):
        sat = np.ascontiguousarray(sat, np.float32)
        out = np.zeros_like(sat)
        self._f("circle")(_f32(sat), sat.shape[1], sat.shape[0], int(r), _f32(out))
        return out

    def blob_list(self, rgba, circ, thr, min_score, radius, max_matches):
        h, w = circ.shape
        m = np.zeros(max(max_matches, 1), MATCH_DTYPE)
        counter = np.zeros(3, np.int32)
        self._f("blob_list")(_u8(np.ascontiguousarray(rgba)), _f32(np.ascontiguousarray(circ, np.float32)), w, h,
                             m.ctypes.data_as(C.c_void_p), counter.ctypes.data_as(C.POINTER(C.c_int32)),
                             C.c_float(thr), C.c_float(min_score), int(radius), int(max_matches))
        return m[: min(int(counter[0]), max_matches)].copy(), counter

    def rgba2nv12(self, rgba):
        h, w = rgba.shape[:2]
        out = np.zeros(2 * w * h, np.uint8)
        self._f("rgba2nv12")(_u8(np.ascontiguousarray(rgba)), w, h, _u8(out))
        return out

    def f2nv12(self, a):
        a = np.ascontiguousarray(a, np.float32)
        h, w = a.shape
        out = np.zeros(2 * w * h, np.uint8)
        self._f("f2nv12")(_f32(a), w, h, _u8(out))
        return out

    def quad2nv12(self, ch, fmt, mode=0):
        hq, wq = ch[0].shape
        out = np.zeros(2 * wq * hq, np.uint8)
        self._f("quad2nv12")(*[_u8(np.ascontiguousarray(c)) for c in ch], fmt, wq, hq, _u8(out), mode)
        return out

    def quad2rgba(self, ch, fmt, mode=0):
        hq, wq = ch[0].shape
        out = np.zeros((hq, wq, 4), np.uint8)
        self._f("quad2rgba")(*[_u8(np.ascontiguousarray(c)) for c in ch], fmt, wq, hq, _u8(out), mode)
        return out

    def circularize(self, a, minr, maxr):
        a = np.ascontiguousarray(a, np.float32)
        out = np.zeros_like(a)
        self._f("circularize")(_f32(a), a.shape[1], a.shape[0], int(minr), int(maxr), _f32(out))
        return out

    def blob_score(self, rgba, circ, thr, radius):
        h, w = circ.shape
        out = np.zeros((h, w), np.float32)
        self._f("blob_score")(_u8(np.ascontiguousarray(rgba)), _f32(np.ascontiguousarray(circ, np.float32)), w, h,
                              C.c_float(thr), int(radius), _f32(out))
        return out

    # ---- whole frame --------------------------------------------------------------------
    def detect(self, raw, p: Params, with_blob_list: bool = True, want_images: bool = True):
        raw = np.ascontiguousarray(raw, np.uint8)
        nf = (p.hf, p.wf)
        out = {}
        if want_images:
            out["flat"] = np.zeros(nf + (4,), np.uint8)
            out["grad"] = np.zeros(nf, np.float32)
            out["sat"] = np.zeros(nf, np.float32)
            out["circ"] = np.zeros(nf, np.float32)
        m = np.zeros(max(p.max_blobs, 1), MATCH_DTYPE)
        counter = np.zeros(3, np.int32)
        mx = self._f("detect")(
            _u8(raw), C.byref(p),
            _u8(out["flat"]) if want_images else None, _f32(out["grad"]) if want_images else None,
            _f32(out["sat"]) if want_images else None, _f32(out["circ"]) if want_images else None,
            m.ctypes.data_as(C.c_void_p), counter.ctypes.data_as(C.POINTER(C.c_int32)), int(with_blob_list))
        out["matches"] = m[: min(int(counter[0]), p.max_blobs)].copy()
        out["counter"] = counter
        out["max_abs_sat"] = float(mx)
        return out


def canonical(matches: np.ndarray) -> np.ndarray:
    """Canonical ordering of a blob list: byte-wise sort of the 22-byte records."""
    raw = np.ascontiguousarray(matches).view(np.uint8).reshape(-1, 22)
    order = np.lexsort(raw.T[::-1])
    return matches[order]
