#!/usr/bin/env python
"""Small driver for ncu: a few steps of the fused path on the headline workload (no CPU oracle, no e2e leg).

  python tools/prof_step.py --batch 8 --group 2 --steps 3 [--times]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vision-processor_b200", "python"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from vpb200 import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--group", type=int, default=0)
ap.add_argument("--lanes", type=int, default=3)
ap.add_argument("--direct", action="store_true", help="direct-gather reprojection instead of the staged kernel")
ap.add_argument("--reproject", type=int, default=2, help="0 direct gather, 2 staged + frame-invariant part hoisted (default)")
ap.add_argument("--chunk", type=int, default=0, help="frames per CTA of the hoisted reprojection (0 = automatic)")
ap.add_argument("--width", type=int, default=2448)
ap.add_argument("--height", type=int, default=2048)
ap.add_argument("--no-gradcirc", action="store_true", help="row sums + streaming circularity instead of the fused gradient + circularity kernel")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--times", action="store_true", help="print CUDA-event time per step and per stage")
args = ap.parse_args()

lp, frames = bench.build_workload(args.width, args.height, 4)
p = lib.params_from_launch(lp)
B, nf, rb = args.batch, lp.wf * lp.hf, frames.shape[1]
dev = torch.device("cuda", 0)
d_raw = torch.empty((B, rb), dtype=torch.uint8, device=dev)
for i in range(B):
    d_raw[i].copy_(torch.from_numpy(frames[i % 4]))
d_flat = torch.empty((B, nf * 4), dtype=torch.uint8, device=dev)
d_grad = torch.empty((B, nf), dtype=torch.float32, device=dev)
d_circ = torch.empty((B, nf), dtype=torch.float32, device=dev)
d_m = torch.zeros((B, p.max_blobs * 22), dtype=torch.uint8, device=dev)
d_c = torch.zeros((B, 3), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
ctx = lib.Context(0)
ctx.set_group(args.group)
ctx.set_lanes(args.lanes)
ctx.set_staged_reproject(0 if args.direct else args.reproject)
ctx.set_hoist_chunk(args.chunk)
ctx.set_fused_gradcirc(not args.no_gradcirc)


def step():
    ctx.detect_batch_device(d_raw.data_ptr(), B, p, d_flat.data_ptr(), d_grad.data_ptr(), d_circ.data_ptr(), d_m.data_ptr(), d_c.data_ptr())


for _ in range(args.warmup):
    step()
ctx.sync()
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record(stream)
for _ in range(args.steps):
    step()
e1.record(stream)
t_issue = time.perf_counter() - t0
e1.synchronize()
ms = e0.elapsed_time(e1)
print(f"batch {B} group {args.group} lanes {args.lanes}: {ms / args.steps / B * 1e3:.2f} us/frame, {B * args.steps / ms * 1e3:.0f} frames/s, "
      f"cpu issue {t_issue / args.steps / B * 1e6:.2f} us/frame, counters[0]={d_c[0].tolist()}")
if args.times:
    ctx.profiling(True)
    step()
    ctx.sync()
    agg = {}
    for name, t in ctx.runtimes():
        a = agg.setdefault(name, [0.0, 0])
        a[0] += t
        a[1] += 1
    for k, (t, n) in agg.items():
        print(f"  {k:14s} {n:4d} launches  {t / B * 1e3:8.2f} us/frame  {t / n * 1e3:8.2f} us/launch")
ctx.close()
