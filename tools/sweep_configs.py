#!/usr/bin/env python
"""BASELINE.json configs 1, 4 and 5 (config 2 is bench.py, config 3 is bench.py --gpus 8):

  config 1  blob_benchmark region (raw2quad .. circle, no blobList) on one 1920x1200 BayerRG8 frame, CPU reference arm
  config 4  full detection + ONE debug-stream NV12 conversion per frame (quad2nv12 / rgb2nv12 / f2nv12 rotated like main.cpp:380-393)
  config 5  batched 4096x3000 frames, batch 1..64, device-resident, vs the HBM roofline

Prints one JSON line per measurement (also appended to profiles/ by the caller).  GPU required except for --config 1."""
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


def gpu_setup(lp, frames, batch):
    import torch
    from vpb200 import lib
    p = lib.params_from_launch(lp)
    nf, rb = lp.wf * lp.hf, frames.shape[1]
    dev = torch.device("cuda", 0)
    t = dict(raw=torch.empty((batch, rb), dtype=torch.uint8, device=dev), flat=torch.empty((batch, nf * 4), dtype=torch.uint8, device=dev),
             grad=torch.empty((batch, nf), dtype=torch.float32, device=dev), circ=torch.empty((batch, nf), dtype=torch.float32, device=dev),
             m=torch.zeros((batch, p.max_blobs * 22), dtype=torch.uint8, device=dev), c=torch.zeros((batch, 3), dtype=torch.int32, device=dev))
    for i in range(batch):
        t["raw"][i].copy_(torch.from_numpy(frames[i % len(frames)]))
    torch.cuda.synchronize()
    return p, t


def timed(ctx, fn, steps, warmup=3):
    import torch
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(warmup):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


def config4(batch=96, steps=10):
    import torch
    from vpb200 import lib
    lp, frames = workload(2448, 2048, 8)
    p, t = gpu_setup(lp, frames, batch)
    nf, nq = lp.wf * lp.hf, lp.wq * lp.hq
    nv12 = torch.empty((batch, 2 * max(nf, nq)), dtype=torch.uint8, device="cuda")
    ctx = lib.Context(0)
    L = ctx.lib

    def detect():
        ctx.detect_batch_device(t["raw"].data_ptr(), batch, p, t["flat"].data_ptr(), t["grad"].data_ptr(), t["circ"].data_ptr(), t["m"].data_ptr(), t["c"].data_ptr())

    def detect_and_stream():
        detect()
        for i in range(batch):  # one view per frame, rotating like main.cpp:380-393: quad, flat, gradDot, blobCenter
            v = i % 4
            out = C.c_void_p(nv12[i].data_ptr())
            if v == 0:
                ctx._ck(L.vp_raw2nv12_device(ctx.h, C.c_void_p(t["raw"][i].data_ptr()), p.fmt, p.wq, p.hq, out, 0))
            elif v == 1:
                ctx._ck(L.vp_rgba2nv12_device(ctx.h, C.c_void_p(t["flat"][i].data_ptr()), p.wf, p.hf, out))
            else:
                ctx._ck(L.vp_f2nv12_device(ctx.h, C.c_void_p((t["grad"] if v == 2 else t["circ"])[i].data_ptr()), p.wf, p.hf, out))

    def detect_and_stream_batched():
        detect()
        q = batch // 4  # one view per frame, the four views over four contiguous quarters of the batch: four launches in all
        stride = nv12.shape[1]
        ctx._ck(L.vp_raw2nv12_batch_device(ctx.h, C.c_void_p(t["raw"][0].data_ptr()), q, p.fmt, p.wq, p.hq, C.c_void_p(nv12[0].data_ptr()), stride, 0))
        ctx._ck(L.vp_rgba2nv12_batch_device(ctx.h, C.c_void_p(t["flat"][q].data_ptr()), q, p.wf, p.hf, C.c_void_p(nv12[q].data_ptr()), stride))
        ctx._ck(L.vp_f2nv12_batch_device(ctx.h, C.c_void_p(t["grad"][2 * q].data_ptr()), q, p.wf, p.hf, C.c_void_p(nv12[2 * q].data_ptr()), stride))
        ctx._ck(L.vp_f2nv12_batch_device(ctx.h, C.c_void_p(t["circ"][3 * q].data_ptr()), batch - 3 * q, p.wf, p.hf, C.c_void_p(nv12[3 * q].data_ptr()), stride))

    ms0 = timed(ctx, detect, steps)
    ms1 = timed(ctx, detect_and_stream, steps)
    ms2 = timed(ctx, detect_and_stream_batched, steps)
    fps0, fps1, fps2 = batch / ms0 * 1e3, batch / ms1 * 1e3, batch / ms2 * 1e3
    print(json.dumps({"config": 4, "workload": "2448x2048 full detection + one NV12 debug-stream conversion per frame (views rotated)", "batch": batch,
                      "frames_per_s_detection_only": fps0, "frames_per_s_with_nv12": fps1, "nv12_us_per_frame": (ms1 - ms0) / batch * 1e3,
                      "frames_per_s_with_nv12_batched": fps2, "nv12_us_per_frame_batched": (ms2 - ms0) / batch * 1e3,
                      "headroom_over_60fps_camera": fps2 / 60.0}))
    ctx.close()


def config5(batches=(1, 2, 4, 8, 16, 32, 64), steps=10):
    from vpb200 import lib
    lp, frames = workload(4096, 3000, 4, n_robots=16, n_balls=4)
    nq, nf = lp.wq * lp.hq, lp.wf * lp.hf
    peak, src = peak_gbs()
    for b in batches:
        p, t = gpu_setup(lp, frames, b)
        ctx = lib.Context(0)

        def detect():
            ctx.detect_batch_device(t["raw"].data_ptr(), b, p, t["flat"].data_ptr(), t["grad"].data_ptr(), t["circ"].data_ptr(), t["m"].data_ptr(), t["c"].data_ptr())

        ms = timed(ctx, detect, steps if b >= 8 else steps * 4)
        blobs = float(np.minimum(t["c"].cpu().numpy()[:, 0], p.max_blobs).mean())
        a_frame = 4 * nq + 12 * nf + 22 * blobs + 12
        fps = b / ms * 1e3
        print(json.dumps({"config": 5, "workload": "4096x3000 BayerRG8 full detection, device-resident", "batch": b, "flat": [lp.wf, lp.hf],
                          "frames_per_s": fps, "us_per_frame": ms / b * 1e3, "algorithmic_bytes_per_frame": a_frame,
                          "hbm_fraction": a_frame * fps / 1e9 / peak, "peak_gbs": peak, "peak_source": src, "circle_radius": lp.circle_radius,
                          "sat_fallbacks": ctx.sat_fallbacks()}))
        ctx.close()
        del t


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, nargs="+", default=[1, 4, 5])
    a = ap.parse_args()
    if 1 in a.config:
        config1()
    if 4 in a.config:
        config4()
    if 5 in a.config:
        config5()
