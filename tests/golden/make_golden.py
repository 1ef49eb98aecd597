#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE kernels (kernel/*.cl compiled in place through oracle/clemu.h ->
oracle/_ref/libvp_clref.so).  Run in the build container, where /root/reference exists:

    make -C oracle && python tests/golden/make_golden.py

Each fixture holds a seeded raw frame, the launch parameters and every intermediate/final output of the reference
pipeline for it.  tests/test_golden.py replays them against the plain-C oracle (CPU) and tests/test_gpu_golden.py
against the CUDA library.  The fixtures are small (tens of KB) and are committed."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "vision-processor_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import common  # noqa: E402
import oracle as O  # noqa: E402

CASES = {
    "rggb_topdown": dict(wq=96, hq=64, fmt=0, seed=1),
    "rggb_tilt_k2": dict(wq=102, hq=66, fmt=0, k2=0.12, tilt=0.2, seed=2),
    "grbg_tilt_k2_trunc": dict(wq=96, hq=64, fmt=1, k2=-0.08, tilt=-0.1, sample_mode=1, seed=3),
    "bgr_nearest": dict(wq=128, hq=96, fmt=2, k2=0.05, sample_mode=2, seed=4),
    "rggb_noise_overflow": dict(wq=80, hq=60, fmt=0, frame="noise", seed=5, max_blobs=40),
}


def params_bytes(p) -> np.ndarray:
    return np.frombuffer(bytes(p), np.uint8).copy()


def main():
    assert O.have_reference(), "oracle/_ref/libvp_clref.so missing: run `make -C oracle` where /root/reference exists"
    ref = O.Oracle("reference")
    for name, kw in CASES.items():
        p, raw, _ = common.make_case(**kw)
        r = ref.detect(raw, p)
        ch = ref.raw2quad(raw, p.fmt, p.wq, p.hq, stale=7)
        hor = ref.sat_horizontal(r["grad"])
        out = dict(
            params=params_bytes(p), raw=raw, flat=r["flat"], grad=r["grad"], hor=hor, sat=r["sat"], circ=r["circ"],
            matches=np.frombuffer(r["matches"].tobytes(), np.uint8), counter=r["counter"], max_abs_sat=np.float64(r["max_abs_sat"]),
            ch0=ch[0], ch1=ch[1], ch2=ch[2], ch3=ch[3],
            quad_rgba=ref.quad2rgba(ch, p.fmt, p.sample_mode),
            nv12_flat=ref.rgba2nv12(r["flat"])[: p.wf * p.hf * 3 // 2],
            nv12_grad=ref.f2nv12(r["grad"])[: p.wf * p.hf * 3 // 2],
            nv12_quad=ref.quad2nv12(ch, p.fmt, p.sample_mode)[: p.wq * p.hq * 3 // 2],
            blob_score=ref.blob_score(r["flat"], r["circ"], p.circ_threshold, p.blob_radius),
            circularize=ref.circularize(r["grad"], 3, 5),
        )
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: flat {p.wf}x{p.hf}, {len(r['matches'])} blobs, counter {r['counter'].tolist()}, {os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
