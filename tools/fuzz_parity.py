#!/usr/bin/env python
"""Randomised parity soak of the fused detection path against the CPU oracle (test infrastructure, like tests/): random sensor
sizes (including flat widths that are not a multiple of 4 and images narrower than one 64-column strip), Bayer order, distortion,
tilt, circle radius 1..13, gradient offset 0..5, threshold, batch sizes 1..9 (odd ones end in half a pair of frames), streams, launch groups and reprojection chunks, every flow
(fused gradient + circularity / row sums), noise and rendered frames.  Prints one line per failing case and a summary.

    python tools/fuzz_parity.py [--cases 200] [--seed 1]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "vision-processor_b200", "python"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p_)

import common  # noqa: E402
import oracle as O  # noqa: E402
from vpb200 import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=200)
ap.add_argument("--seed", type=int, default=1)
args = ap.parse_args()
rng = np.random.default_rng(args.seed)
port = O.Oracle("port")
port.set_threads(os.cpu_count() or 1)
fails, plans = 0, {}
with lib.Context(0) as ctx:
    for case in range(args.cases):
        wq = int(rng.integers(20, 260))
        hq = int(rng.integers(16, 160))
        fmt = int(rng.integers(0, 2))
        k2 = float(rng.choice([0.0, 0.05, 0.12]))
        tilt = float(rng.choice([0.0, 0.0, 0.15]))
        radius = int(rng.integers(1, 14))
        offset = int(rng.integers(0, 6))
        n = int(rng.integers(1, 10))
        flow = int(rng.choice([2, 2, 0]))
        scale = float(rng.choice([3.0, 4.0, 5.5]))
        kind = str(rng.choice(["scene", "scene", "noise"]))
        lanes = int(rng.choice([1, 3, 3]))
        group = int(rng.choice([0, 0, int(rng.integers(1, n + 1))]))
        chunk = int(rng.choice([0, 4, 8]))
        desc = dict(case=case, wq=wq, hq=hq, fmt=fmt, k2=k2, tilt=tilt, radius=radius, offset=offset, n=n, flow=flow, scale=scale, kind=kind,
                    lanes=lanes, group=group, chunk=chunk)
        try:
            frames = []
            for i in range(n):
                p, raw, _ = common.make_case(wq=wq, hq=hq, fmt=fmt, k2=k2, tilt=tilt, seed=1000 * case + i, scale_mm=scale, n_robots=2, n_balls=1,
                                             thr=0.0 if radius == 1 else 6.0, frame=kind)
                frames.append(raw)
            p.circle_radius, p.blob_radius, p.grad_offset = radius, max(radius - 1, 0), offset
            desc["flat"] = (p.wf, p.hf)
            vp = common.to_vp(p)
            ctx.set_fused_gradcirc(flow)
            ctx.set_lanes(lanes)
            ctx.set_group(group)
            ctx.set_hoist_chunk(chunk)
            got = common.detect_device(ctx, frames, vp)
            rp = got["plan"]["reproject"]
            plans[f"reproject {rp}"] = plans.get(f"reproject {rp}", 0) + 1
            plans[got["plan"]["circ"]] = plans.get(got["plan"]["circ"], 0) + 1
            for i in sorted(set([0, n - 1, n // 2])):
                common.assert_frame_equal(got, i, port.detect(frames[i], p))
            if n <= 2:  # the host API's lone-frame path as well
                one = ctx.detect(frames[0], vp)
                w = port.detect(frames[0], p)
                np.testing.assert_array_equal(one["flat"], w["flat"])
                np.testing.assert_array_equal(one["counter"][0], w["counter"])
                common.assert_matches_equal(one["matches"][0], w["matches"])
        except Exception as e:  # noqa: BLE001
            fails += 1
            print("FAIL", desc, type(e).__name__, str(e).splitlines()[0][:200], flush=True)
        finally:
            ctx.set_fused_gradcirc(1)
            ctx.set_lanes(3)
            ctx.set_group(0)
            ctx.set_hoist_chunk(0)
print(f"{args.cases} cases, {fails} failures; circularity flows taken: {plans}")
sys.exit(1 if fails else 0)
