#!/usr/bin/env python
"""Lone-frame latency of vp_detect_host (p50/p99 over 1000 frames from a pinned ring) for the two circularity flows and a few
segment heights: python tools/latency_ab.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vision-processor_b200", "python"))
import bench  # noqa: E402
from vpb200 import lib  # noqa: E402

lp, frames = bench.build_workload(2448, 2048, 8)
p = lib.params_from_launch(lp)
rb = frames.shape[1]
ring = lib.PinnedArray((8, rb), np.uint8)
ring.array[:] = frames
pm = lib.PinnedArray((p.max_blobs * 22,), np.uint8)
pc = lib.PinnedArray((1, 3), np.int32)
for gc in (2, 0):  # 2: the fused gradient + circularity kernel also on lone frames; 0: the row-sum flow (what lone frames take by default)
    for seg in (os.environ.get("VP_CIRC_SEG", "auto"),):
        ctx = lib.Context(0)
        ctx.set_fused_gradcirc(gc)
        lat = []
        for i in range(1040):
            t0 = time.perf_counter()
            ctx.detect_host_into(ring.ptr.value + (i % 8) * rb, 1, p, pm.ptr.value, pc.ptr.value)
            lat.append(1e3 * (time.perf_counter() - t0))
        lat = np.sort(np.array(lat[40:]))
        print(f"fused_gradcirc={gc} seg={seg} plan={ctx.last_plan()} p50 {lat[len(lat) // 2]:.4f} ms  p99 {lat[int(len(lat) * 0.99)]:.4f} ms  counters {pc.array[0].tolist()}")
        ctx.close()
