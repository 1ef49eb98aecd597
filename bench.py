#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native vision-processor detection path.

Metric (BASELINE.json): 2448x2048 BayerRG8 frames/s, full detection pipeline
(demosaic -> reproject -> gradientDot -> row sums -> circularity -> blob list), aggregate over N GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]

One "step" = one pass of the fused path (vp_detect_batch_device) over a batch of B synthetic frames that are
already resident in HBM.  B frames x 5.0 MB = more than the 126 MB L2, so every step streams its inputs from
HBM.  `e2e` is the same metric through vp_detect_host: pinned HOST frames in, blob lists + counters back in
host memory, copies inside the timed region.  Prints ONE JSON line on rank 0.

The CPU oracle (oracle/) is executed only by the `cpu_baseline` leg and by `--impl reference`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "vision-processor_b200", "python"))

from vpb200 import geometry as G, synth as S  # noqa: E402

SENSOR_W, SENSOR_H = 2448, 2048
METRIC = "2448x2048 Bayer frames/sec (full detection pipeline, aggregate over GPUs)"
UNIT = "frames/s"
WORKLOAD = "single camera 2448x2048 BayerRG8 full detection pipeline (BASELINE configs[1]); one camera stream per GPU"


# ------------------------------------------------------------------------------------------------ workload
def build_workload(n_distinct: int, cam_seed: int = 0):
    """The synthetic camera of SURVEY 8(d) + n_distinct noisy frames of one rendered SSL scene."""
    wq, hq = SENSOR_W // 2, SENSOR_H // 2
    cam = G.default_camera(wq, hq, k2=0.0)
    persp = G.Perspective(cam)
    persp.geometry_check(wq, hq, 180.0)
    lp = G.launch_params(persp, S.FMT_RGGB, wq, hq)
    scene = S.random_scene(persp.visible_field_extent, 16, 4, seed=1 + cam_seed)
    clean = S.render_rgb(scene, cam, SENSOR_W, SENSOR_H)
    frames = np.stack([S.render_raw(scene, cam, SENSOR_W, SENSOR_H, S.FMT_RGGB, seed=1000 * cam_seed + i, clean_rgb=clean).reshape(-1)
                       for i in range(n_distinct)])
    return lp, frames


def algorithmic_bytes(lp, blobs_per_frame: float) -> dict:
    """compulsory input plus the outputs the boundary returns, per frame; and per stage."""
    nq, nf = lp.wq * lp.hq, lp.wf * lp.hf
    return {
        "frame": 4 * nq + 4 * nf + 4 * nf + 4 * nf + 22 * blobs_per_frame + 12,
        "reproject": 4 * nq + 4 * nf,          # raw in, flat out
        "grad_sat": 4 * nf + 8 * nf,           # (single-pass alternative) flat in, gradDot + SAT out
        "grad_rowscan": 4 * nf + 8 * nf,       # flat in, gradDot + row prefix sums out
        "colscan": 8 * nf,                     # (SAT-based alternative) row sums in, SAT out
        "circ_peaks": 8 * nf,                  # row sums (or SAT) in, blobCenter out (+ blob bit masks)
        "peaks_emit": 22 * blobs_per_frame,    # sparse: records out
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_pipeline(kind_pref: str = "reference"):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    kind = "reference" if (kind_pref == "reference" and O.have_reference()) else "port"
    return O, O.Oracle(kind), kind


def oracle_params(O, lp):
    import ctypes as C
    p = O.Params()
    p.fmt, p.wq, p.hq, p.wf, p.hf = lp.fmt, lp.wq, lp.hq, lp.wf, lp.hf
    C.memmove(C.byref(p.model), lp.model_bytes, 72)
    p.max_robot_height, p.field_scale, p.off_x, p.off_y = lp.max_robot_height, lp.field_scale, lp.off_x, lp.off_y
    p.grad_offset, p.circle_radius, p.circ_threshold, p.min_score = lp.grad_offset, lp.circle_radius, lp.circ_threshold, lp.min_score
    p.blob_radius, p.max_blobs, p.sample_mode = lp.blob_radius, lp.max_blobs, lp.sample_mode
    return p


def time_cpu(orc, p, frames, n_frames: int) -> float:
    t0 = time.perf_counter()
    for i in range(n_frames):
        orc.detect(frames[i % len(frames)], p, with_blob_list=True, want_images=False)
    return time.perf_counter() - t0


def cpu_baseline(lp, frames, budget_s: float = 12.0) -> dict:
    O, orc, kind = cpu_pipeline()
    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    p = oracle_params(O, lp)
    t1 = time_cpu(orc, p, frames, 1)  # warm-up + calibration
    n = int(max(2, min(60, budget_s / max(t1, 1e-3))))
    dt = time_cpu(orc, p, frames, n)
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} frames of the same 2448x2048 workload, all stages incl. blobList, {cores} OpenMP threads, "
                      f"{'reference kernel/*.cl compiled in place through oracle/clemu.h' if kind == 'reference' else 'oracle/vp_oracle.c restatement'}"}


def run_reference(args, rank: int):
    if rank != 0:
        return
    lp, frames = build_workload(4)
    O, orc, kind = cpu_pipeline()
    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    p = oracle_params(O, lp)
    t1 = time_cpu(orc, p, frames, 1)
    per_step = int(max(1, min(32, 3.0 / max(t1, 1e-3))))  # ~3 s of CPU work per step
    for _ in range(args.warmup):
        time_cpu(orc, p, frames, per_step)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_cpu(orc, p, frames, per_step)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} frames per step ({kind}: " + ("reference kernel/*.cl compiled in place via oracle/clemu.h" if kind == "reference" else "oracle/vp_oracle.c") + f"), {cores} OpenMP threads"
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "frames_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ B200 arm
_RESULT_FD = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version banner at the first
    collective, for one) is sent to stderr by pointing fd 1 at fd 2; the result goes to the saved descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, line)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=192, help="frames per step and per GPU (three groups of 64 frames on three streams)")
    ap.add_argument("--e2e-batch", type=int, default=32)
    ap.add_argument("--group", type=int, default=0, help="frames per launch group (0 = automatic)")
    ap.add_argument("--lanes", type=int, default=0, help="streams the groups are spread over (0 = library default)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from vpb200 import lib

    numa_cpus = []
    if world > 1:  # one camera stream per GPU: keep this rank's pinned frame ring on the socket its GPU hangs off
        from vpb200 import shard
        numa_cpus = shard.bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload: camera `rank` of the field, B frames per step ---------------------------------
    n_distinct = 8
    lp, frames = build_workload(n_distinct, cam_seed=rank)
    p = lib.params_from_launch(lp)
    B, nf, rb = args.batch, lp.wf * lp.hf, frames.shape[1]
    dev = torch.device("cuda", local_rank)
    d_raw = torch.empty((B, rb), dtype=torch.uint8, device=dev)
    h_frames = torch.from_numpy(frames)
    for i in range(B):
        d_raw[i].copy_(h_frames[i % n_distinct])
    d_flat = torch.empty((B, nf * 4), dtype=torch.uint8, device=dev)
    d_grad = torch.empty((B, nf), dtype=torch.float32, device=dev)
    d_circ = torch.empty((B, nf), dtype=torch.float32, device=dev)
    d_matches = torch.zeros((B, p.max_blobs * 22), dtype=torch.uint8, device=dev)
    d_counter = torch.zeros((B, 3), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    ctx = lib.Context(local_rank)
    ctx.set_group(args.group)
    if args.lanes:
        ctx.set_lanes(args.lanes)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step():
        ctx.detect_batch_device(d_raw.data_ptr(), B, p, d_flat.data_ptr(), d_grad.data_ptr(), d_circ.data_ptr(), d_matches.data_ptr(), d_counter.data_ptr())

    for _ in range(args.warmup):
        step()
    ctx.sync()
    counters = d_counter.cpu().numpy()
    assert (counters[:, 0] > 0).all() and ctx.sat_fallbacks() == 0, "warm-up produced no blobs or left the exact SAT range"
    for i in range(n_distinct, B):  # identical frames must give identical counters
        assert (counters[i] == counters[i % n_distinct]).all()
    blobs_per_frame = float(np.minimum(counters[:, 0], p.max_blobs).mean())

    # ---- timed region: K steps, device-resident inputs -------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    e1.synchronize()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    barrier()
    launches = ctx.launch_count() - launches0
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- same K steps again with a CUDA event pair around every kernel (per-kernel durations) ----
    ctx.profiling(True)
    step()
    ctx.sync()
    first = ctx.runtimes()
    n_prof_steps = max(1, min(args.steps, 4000 // max(1, len(first))))
    for _ in range(n_prof_steps):
        step()
    ctx.sync()
    per_stage = {}
    for name, ms in ctx.runtimes():
        s_ = per_stage.setdefault(name, [0.0, 0])
        s_[0] += ms
        s_[1] += 1
    ctx.profiling(False)

    abytes = algorithmic_bytes(lp, blobs_per_frame)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_gbs, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    except Exception:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    total_ms = sum(v[0] for v in per_stage.values()) or 1.0
    top = max((k for k in per_stage if k in abytes), key=lambda k: per_stage[k][0])
    frames_per_launch = B * n_prof_steps / per_stage[top][1]  # frames of the profiled pass / launches of that kernel
    avg_ms = per_stage[top][0] / per_stage[top][1]
    achieved = abytes[top] * frames_per_launch / (avg_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per frame of each kernel, from the committed ncu --set full capture
    # (profiles/ncu_traffic.json, written by tools/ncu_traffic.py), scaled to the frames one launch of this run processes
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
        traffic = ncu[top]["dram_bytes_per_frame"] * frames_per_launch if top in ncu else None
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes[top] * frames_per_launch,
                "avg_launch_ms": avg_ms, "share_of_step": per_stage[top][0] / total_ms}
    pipeline_gbs = abytes["frame"] * (value / world) / 1e9
    stage_ms = {k: {"ms_per_frame": v[0] / (B * n_prof_steps), "share": v[0] / total_ms} for k, v in per_stage.items()}

    # ---- e2e: host frames in, blob lists out, through vp_detect_host ------------------------------
    Be = args.e2e_batch
    pin_raw = lib.PinnedArray((Be, rb), np.uint8)
    pin_m = lib.PinnedArray((Be, p.max_blobs * 22), np.uint8)
    pin_c = lib.PinnedArray((Be, 3), np.int32)
    for i in range(Be):
        pin_raw.array[i] = frames[i % n_distinct]

    def e2e_step():
        ctx.detect_host_into(pin_raw.ptr.value, Be, p, pin_m.ptr.value, pin_c.ptr.value)

    for _ in range(3):
        e2e_step()
    assert (pin_c.array[:n_distinct] == counters[:n_distinct]).all(), "host path and device path disagree"
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * Be * e2e_steps / e2e_s

    # ---- single-frame latency through the host API (p50 / p99) ------------------------------------
    lat = []
    replays0 = ctx.latency_graph_replays()
    for i in range(300):
        t0 = time.perf_counter()
        ctx.detect_host_into(pin_raw.ptr.value + (i % Be) * rb, 1, p, pin_m.ptr.value, pin_c.ptr.value)
        lat.append(1e3 * (time.perf_counter() - t0))
    lat = np.sort(np.array(lat[20:]))
    latency = {"p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99)]), "frames": int(len(lat)),
               "path": "vp_detect_host, 1 frame per call, pinned host in -> host blob list out; upload in 2 chunks of rows with the "
                       "reprojection of the first under the second, one download, replayed as a CUDA graph",
               "graph_replays": int(ctx.latency_graph_replays() - replays0)}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32/f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "frames_per_step_per_gpu": B, "flat_size": [lp.wf, lp.hf], "max_blobs": p.max_blobs, "blobs_per_frame": blobs_per_frame,
                   "l2": f"inputs larger than L2: {B} frames x {rb} B = {B * rb / 1e6:.0f} MB raw per step, streamed from HBM every step",
                   "parallelism": f"{world} independent camera streams, no collective",
                   "host_cpus_of_rank0": len(numa_cpus) if numa_cpus else "all"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": Be * rb, "d2h_bytes_per_step": Be * (p.max_blobs * 22 + 12),
                "frames_per_step": Be, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_pipeline": {"bound": "hbm", "achieved": pipeline_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": pipeline_gbs / peak_gbs,
                              "algorithmic_bytes_per_frame": abytes["frame"], "note": "per GPU; compulsory input + API outputs per frame x frames/s"},
        "stage_ms": stage_ms,
        "latency": latency,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(lp, frames)
    if rank == 0:
        emit(out)
    pin_raw.free(); pin_m.free(); pin_c.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
