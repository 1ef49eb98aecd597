#!/usr/bin/env python
"""BASELINE.json config 1 -- the reference's own CPU-runnable case: the timed region of blob_benchmark (raw2quad .. circle, no
blobList) on one 1920x1200 BayerRG8 frame, with the reference kernels compiled in place (oracle/_ref), 1 thread and all threads.

Configs 2, 4 and 5 (and 3 = --gpus N) come out of bench.py itself: bench.py --config {2,4,5} [--frame-size WxH] [--batch N].
Prints one JSON line per measurement.  No GPU needed."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vision-processor_b200", "python"))

from vpb200 import geometry as G, synth as S  # noqa: E402


def workload(sensor_w, sensor_h, n_distinct, n_robots=16, n_balls=4):
    wq, hq = sensor_w // 2, sensor_h // 2
    cam = G.default_camera(wq, hq, k2=0.0)
    persp = G.Perspective(cam)
    persp.geometry_check(wq, hq, 180.0)
    lp = G.launch_params(persp, S.FMT_RGGB, wq, hq)
    scene = S.random_scene(persp.visible_field_extent, n_robots, n_balls, seed=1)
    clean = S.render_rgb(scene, cam, sensor_w, sensor_h)
    frames = np.stack([S.render_raw(scene, cam, sensor_w, sensor_h, S.FMT_RGGB, seed=i, clean_rgb=clean).reshape(-1) for i in range(n_distinct)])
    return lp, frames


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def config1():
    """CPU only: the timed region of blob_benchmark.cpp:143-158 with the reference kernels compiled in place."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bench
    import oracle as O
    lp, frames = workload(1920, 1200, 1, n_robots=8, n_balls=2)
    kind = "reference" if O.have_reference() else "port"
    orc = O.Oracle(kind)
    p = bench.oracle_params(O, lp)
    for threads in (1, os.cpu_count() or 1):
        orc.set_threads(threads)
        orc.detect(frames[0], p, with_blob_list=False, want_images=False)
        n = 10 if threads == 1 else 40
        t0 = time.perf_counter()
        for _ in range(n):
            orc.detect(frames[0], p, with_blob_list=False, want_images=False)
        dt = (time.perf_counter() - t0) / n
        print(json.dumps({"config": 1, "workload": "blob_benchmark region on one 1920x1200 BayerRG8 frame (raw2quad..circle)", "impl": f"cpu {kind}",
                          "threads": threads, "ms_per_frame": 1e3 * dt, "frames_per_s": 1 / dt, "flat": [lp.wf, lp.hf]}))




if __name__ == "__main__":
    config1()
