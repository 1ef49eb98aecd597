#!/usr/bin/env python
"""Single-frame latency of vp_detect_host on the headline frame (2448x2048) as a function of the number of upload
strips: pinned host frame in -> host blob list out, wall clock around the blocking call.  Prints one line per setting.

  python tools/latency_sweep.py [--strips 1,2,4,6,8,12] [--frames 300]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vision-processor_b200", "python"))

import bench  # noqa: E402
from vpb200 import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--strips", default="1,2,4,6,8,12")
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--no-graph", action="store_true", help="direct launches instead of the CUDA graph replay")
ap.add_argument("--ring", type=int, default=4, help="pinned frame buffers the calls rotate through (1 = always the same address)")
args = ap.parse_args()

lp, frames = bench.build_workload(2448, 2048, 4)
args.ring = max(1, min(args.ring, 32))
p = lib.params_from_launch(lp)
rb = p.raw_frame_bytes()
pin_raw = lib.PinnedArray((args.ring, rb), np.uint8)
for i in range(args.ring):
    pin_raw.array[i] = frames[i % 4]
pin_m = lib.PinnedArray((p.max_blobs * 22,), np.uint8)
pin_c = lib.PinnedArray((1, 3), np.int32)

with lib.Context(0) as ctx:
    ref = None
    ctx.set_latency_graph(not args.no_graph)
    for n in [int(x) for x in args.strips.split(",")]:
        ctx.set_strips(n)
        lat = []
        r0 = ctx.latency_graph_replays()
        for i in range(args.frames + 20):
            t0 = time.perf_counter()
            ctx.detect_host_into(pin_raw.ptr.value + (i % args.ring) * rb, 1, p, pin_m.ptr.value, pin_c.ptr.value)
            lat.append(1e3 * (time.perf_counter() - t0))
        lat = np.sort(np.array(lat[20:]))
        got = (pin_c.array.copy().tolist(), pin_m.array[: 22 * min(int(pin_c.array[0, 0]), p.max_blobs)].tobytes())
        if ref is None:
            ref = got
        same = got == ref
        print(json.dumps({"strips": n, "p50_ms": round(float(lat[len(lat) // 2]), 4), "p99_ms": round(float(lat[int(len(lat) * 0.99)]), 4),
                          "min_ms": round(float(lat[0]), 4), "counter": got[0], "same_as_first": same, "graph_replays": ctx.latency_graph_replays() - r0, "ring": args.ring}), flush=True)
