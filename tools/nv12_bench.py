#!/usr/bin/env python
"""Debug-stream conversions (rgba2nv12.cl, f2nv12.cl, quad2nv12.cl) as batched device-to-device launches: us per view and
fraction of the HBM roofline.  Algorithmic bytes per view: 4 B/px in + 1.5 B/px out.

  python tools/nv12_bench.py [--batch 64] [--steps 20]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vision-processor_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from vpb200 import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--wq", type=int, default=1224)
ap.add_argument("--hq", type=int, default=1024)
args = ap.parse_args()

peak = 6550.1
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

B, w, h = args.batch, args.wq, args.hq
n = w * h
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(1)
rgba = torch.randint(0, 256, (B, n * 4), dtype=torch.uint8, generator=g).to(dev)
f32 = (torch.rand((B, n), generator=g) * 400 - 200).to(dev)
raw = torch.randint(0, 256, (B, n * 4), dtype=torch.uint8, generator=g).to(dev)  # 2448x2048 Bayer bytes
stride = 2 * n
out = torch.empty((B, stride), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()

with lib.Context(0) as ctx:
    L = ctx.lib
    stream = torch.cuda.ExternalStream(ctx.stream)
    calls = {
        "rgba2nv12": lambda: L.vp_rgba2nv12_batch_device(ctx.h, C.c_void_p(rgba.data_ptr()), B, w, h, C.c_void_p(out.data_ptr()), stride),
        "f2nv12": lambda: L.vp_f2nv12_batch_device(ctx.h, C.c_void_p(f32.data_ptr()), B, w, h, C.c_void_p(out.data_ptr()), stride),
        "quad2nv12 (from raw)": lambda: L.vp_raw2nv12_batch_device(ctx.h, C.c_void_p(raw.data_ptr()), B, 0, w, h, C.c_void_p(out.data_ptr()), stride, 0),
    }
    for name, fn in calls.items():
        for _ in range(3):
            ctx._ck(fn())
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            ctx._ck(fn())
        e1.record(stream)
        e1.synchronize()
        us = e0.elapsed_time(e1) / args.steps / B * 1e3
        gbs = 5.5 * n / us / 1e3
        print(f"{name:22s} {us:7.3f} us/view  {gbs:7.1f} GB/s algorithmic  {gbs / peak:5.2f} of {peak:.0f} GB/s   ({B} views of {w}x{h} per launch, "
              f"{B * 5.5 * n / 1e6:.0f} MB per launch)", flush=True)
