#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native vision-processor detection path.

Metric (BASELINE.json): Bayer frames/s through the full detection pipeline
(demosaic -> reproject -> gradientDot -> box sums -> circularity -> blob list), aggregate over N GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|4|5] [--frame-size WxH] [--batch B]
                  [--impl b200|reference] [--copy-only]

  --config 2 (default)  BASELINE configs[1]/[2]: one 2448x2048 BayerRG8 camera per GPU, full detection
  --config 4            config 2 + one NV12 debug-stream view per frame (the four views of main.cpp:380-393 rotated)
  --config 5            4096x3000 BayerRG8, batched (use --batch 1..64 for the sweep of BASELINE configs[4])
  --frame-size WxH      any other sensor size (BayerRG8), same camera model

One "step" = `passes` passes of the fused path (vp_detect_batch_device) over a ring of B synthetic frames that are
already resident in HBM (B frames x 5.0 MB = 963 MB, more than the 126 MB L2: every pass streams its inputs from HBM);
`passes` is chosen once, before the timed region, so that the K timed steps last at least a second -- it is stated in
the JSON line (`run.passes_per_step`), and `value` counts every frame of every pass.  `e2e` is the same metric through
vp_detect_host: pinned HOST frames in, blob lists + counters back in host memory, copies inside the timed region, again
at least a second long.  Prints ONE JSON line on rank 0.

--copy-only runs the host-fed pattern of `e2e` (same pinned buffers, same chunking, same streams) with NO kernels and
reports GB/s per rank: the ceiling the host can feed, to set beside the e2e figure at N = 1/2/4/8.

The CPU oracle (oracle/) is executed only by the `cpu_baseline` leg and by `--impl reference`.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "vision-processor_b200", "python"))

from vpb200 import geometry as G, synth as S  # noqa: E402

UNIT = "frames/s"
CONFIGS = {
    2: dict(size=(2448, 2048), batch=384, what="single camera 2448x2048 BayerRG8 full detection pipeline (BASELINE configs[1]); one camera stream per GPU"),
    4: dict(size=(2448, 2048), batch=384, what="2448x2048 BayerRG8 full detection + one NV12 debug-stream view per frame, views rotated as main.cpp:380-393 "
                                               "(BASELINE configs[3]); one camera stream per GPU"),
    5: dict(size=(4096, 3000), batch=64, what="batched 4096x3000 BayerRG8 full detection (BASELINE configs[4]); one batch stream per GPU"),
}


def metric_name(w, h):
    return f"{w}x{h} Bayer frames/sec (full detection pipeline, aggregate over GPUs)"


# ------------------------------------------------------------------------------------------------ workload
def build_workload(sensor_w: int, sensor_h: int, n_distinct: int, cam_seed: int = 0, k2: float = 0.0, tilt: float = 0.0):
    """The synthetic camera of SURVEY 8(d) + n_distinct noisy frames of one rendered SSL scene.  k2 / tilt: the stress camera of
    SURVEY 8(d) (radial distortion, rotation about the x axis in rad) instead of the undistorted top-down one."""
    wq, hq = sensor_w // 2, sensor_h // 2
    cam = G.default_camera(wq, hq, k2=k2)
    if tilt:
        c, s_ = math.cos(tilt / 2), math.sin(tilt / 2)
        cam = G.CameraModel(size=(wq, hq), focal_length=float(wq), principal_point=(wq / 2.0, hq / 2.0), distortion_k2=k2,
                            pos=(0.0, 0.0, 5000.0), quat_wxyz=(s_, -c, 0.0, 0.0))
    persp = G.Perspective(cam)
    persp.geometry_check(wq, hq, 180.0)
    lp = G.launch_params(persp, S.FMT_RGGB, wq, hq)
    scene = S.random_scene(persp.visible_field_extent, 16, 4, seed=1 + cam_seed)
    clean = S.render_rgb(scene, cam, sensor_w, sensor_h)
    frames = np.stack([S.render_raw(scene, cam, sensor_w, sensor_h, S.FMT_RGGB, seed=1000 * cam_seed + i, clean_rgb=clean).reshape(-1)
                       for i in range(n_distinct)])
    return lp, frames


def workload_config(args, lp, world: int) -> dict:
    """`config` of the JSON line: the workload only, identical (keys AND values) in the b200 and the reference arm."""
    w, h = args.size
    rb = w * h
    return {"workload": CONFIGS[args.config]["what"] if not args.frame_size else f"{w}x{h} BayerRG8 full detection pipeline; one stream per GPU",
            "config_id": args.config, "frame_size": [w, h], "flat_size": [lp.wf, lp.hf], "max_blobs": lp.max_blobs,
            "circle_radius": lp.circle_radius, "grad_offset": lp.grad_offset,
            "camera": {"k2": args.k2, "tilt_rad": args.tilt},
            "frames_in_ring_per_gpu": args.batch,
            "l2": f"inputs larger than L2: a ring of {args.batch} frames x {rb} B = {args.batch * rb / 1e6:.0f} MB raw per GPU, streamed from HBM on every pass"
                  if args.batch * rb > 126e6 else
                  f"ring of {args.batch} frames x {rb} B = {args.batch * rb / 1e6:.0f} MB raw: smaller than L2; 256 MB are written between timed steps to flush it",
            "parallelism": f"{world} independent camera streams, no collective"}


def algorithmic_bytes(lp, blobs_per_frame: float) -> dict:
    """compulsory input plus the outputs the boundary returns, per frame; and per stage."""
    nq, nf = lp.wq * lp.hq, lp.wf * lp.hf
    return {
        "frame": 4 * nq + 4 * nf + 4 * nf + 4 * nf + 22 * blobs_per_frame + 12,
        "reproject": 4 * nq + 4 * nf,          # raw in, flat out
        "grad_circ": 4 * nf + 8 * nf,          # flat in, gradDot + blobCenter out (fused gradient + circularity kernel)
        "grad_sat": 4 * nf + 8 * nf,           # (single-pass alternative) flat in, gradDot + SAT out
        "grad_rowscan": 4 * nf + 8 * nf,       # (alternative) flat in, gradDot + row prefix sums out
        "colscan": 8 * nf,                     # (SAT-based alternative) row sums in, SAT out
        "circ_peaks": 8 * nf,                  # (alternative) row sums (or SAT) in, blobCenter out (+ blob bit masks)
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


def time_cpu(orc, p, frames, n_frames: int, nv12: bool = False) -> float:
    t0 = time.perf_counter()
    for i in range(n_frames):
        r = orc.detect(frames[i % len(frames)], p, with_blob_list=True, want_images=nv12)
        if nv12:  # config 4: one debug view per frame, rotated
            v = i % 4
            if v == 0:
                orc.raw2nv12(frames[i % len(frames)], p.fmt, p.wq, p.hq) if hasattr(orc, "raw2nv12") else orc.rgba2nv12(r["flat"])
            elif v == 1:
                orc.rgba2nv12(r["flat"])
            else:
                orc.f2nv12(r["grad"] if v == 2 else r["circ"])
    return time.perf_counter() - t0


def cpu_sample_text(kind: str, cores: int) -> str:
    return ("reference kernel/*.cl compiled in place through oracle/clemu.h" if kind == "reference" else "oracle/vp_oracle.c restatement") + f", {cores} OpenMP threads"


def cpu_baseline(args, lp, frames, budget_s: float = 12.0) -> dict:
    O, orc, kind = cpu_pipeline()
    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    p = oracle_params(O, lp)
    nv12 = args.config == 4
    t1 = time_cpu(orc, p, frames, 1, nv12)  # warm-up + calibration
    n = int(max(2, min(60, budget_s / max(t1, 1e-3))))
    dt = time_cpu(orc, p, frames, n, nv12)
    w, h = args.size
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} frames of the same {w}x{h} workload, all stages incl. blobList, " + cpu_sample_text(kind, cores)}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    lp, frames = build_workload(*args.size, 4, k2=args.k2, tilt=args.tilt)
    O, orc, kind = cpu_pipeline()
    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    p = oracle_params(O, lp)
    nv12 = args.config == 4
    t1 = time_cpu(orc, p, frames, 1, nv12)
    per_step = int(max(1, min(32, 3.0 / max(t1, 1e-3))))  # ~3 s of CPU work per step
    for _ in range(args.warmup):
        time_cpu(orc, p, frames, per_step, nv12)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_cpu(orc, p, frames, per_step, nv12)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} frames per step, " + cpu_sample_text(kind, cores)
    emit(({
        "impl": "reference", "metric": metric_name(*args.size), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f32",
        "data": "synthetic", "config": workload_config(args, lp, world),
        "run": {"frames_per_step": per_step, "note": "bounded sample of the same workload per step, host cores only"},
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


def traffic_from_capture(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per frame of `kernel` from the committed ncu --set full capture
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py): NOT measured in this run -- ncu cannot run inside a timed
    benchmark -- so it is labelled with where it came from and dropped when the capture does not know the kernel."""
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        k = cap["kernels"].get(kernel)
        if not k:
            return None, f"no committed ncu capture of '{kernel}'"
        return k["dram_bytes_per_frame"], f"committed ncu --set full capture ({cap.get('source', 'profiles/ncu_traffic.json')}, {cap.get('frames_per_launch', '?')} frames per launch), scaled to this run's frames per launch; not measured live"
    except Exception as e:  # noqa: BLE001
        return None, f"profiles/ncu_traffic.json unreadable: {e}"


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--frame-size", default="", help="WxH of the Bayer sensor (overrides the size of --config)")
    ap.add_argument("--k2", type=float, default=0.0, help="radial distortion of the synthetic camera (SURVEY 8(d) stress case: 0.12)")
    ap.add_argument("--tilt", type=float, default=0.0, help="rotation of the synthetic camera about the x axis in rad (stress case: 0.2)")
    ap.add_argument("--batch", type=int, default=0, help="frames in the device-resident ring per GPU (0 = the config's default: three groups of 128 frames on three streams)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="the K timed steps last at least this long (passes per step are scaled up)")
    ap.add_argument("--e2e-batch", type=int, default=32)
    ap.add_argument("--group", type=int, default=0, help="frames per launch group (0 = automatic)")
    ap.add_argument("--lanes", type=int, default=0, help="streams the groups are spread over (0 = library default)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--copy-only", action="store_true", help="the pinned H2D + D2H pattern of e2e with no kernels: GB/s per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    args.size = tuple(int(v) for v in args.frame_size.lower().split("x")) if args.frame_size else CONFIGS[args.config]["size"]
    if not args.batch:
        args.batch = CONFIGS[args.config]["batch"]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
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

    def gather_ranks(x: float) -> list:
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    # ---- workload: camera `rank` of the field, a ring of B frames ---------------------------------
    n_distinct = 8 if args.size[0] * args.size[1] <= 6e6 else 4
    lp, frames = build_workload(*args.size, n_distinct, cam_seed=rank, k2=args.k2, tilt=args.tilt)
    p = lib.params_from_launch(lp)
    B, nf, nq, rb = args.batch, lp.wf * lp.hf, lp.wq * lp.hq, frames.shape[1]
    dev = torch.device("cuda", local_rank)
    ctx = lib.Context(local_rank)
    ctx.set_group(args.group)
    if args.lanes:
        ctx.set_lanes(args.lanes)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- host-fed buffers (e2e, copy-only, latency) -------------------------------------------------
    Be = args.e2e_batch
    pin_raw = lib.PinnedArray((Be, rb), np.uint8)
    pin_m = lib.PinnedArray((Be, p.max_blobs * 22), np.uint8)
    pin_c = lib.PinnedArray((Be, 3), np.int32)
    for i in range(Be):
        pin_raw.array[i] = frames[i % n_distinct]

    if args.copy_only:
        # the copy pattern of vp_detect_host (chunks of 4 frames up on one stream, records + counters down on another, three
        # device slots in flight) with no kernels in between: what the host side of this box can feed
        chunk = 4
        slots = [dict(raw=torch.empty((chunk, rb), dtype=torch.uint8, device=dev), m=torch.empty((chunk, p.max_blobs * 22), dtype=torch.uint8, device=dev),
                      c=torch.empty((chunk, 3), dtype=torch.int32, device=dev)) for _ in range(3)]
        h_raw = torch.from_numpy(pin_raw.array)
        h_m, h_c = torch.from_numpy(pin_m.array), torch.from_numpy(pin_c.array)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def copy_pass():
            evs = []
            for k, f0 in enumerate(range(0, Be, chunk)):
                sl = slots[k % 3]
                g = min(chunk, Be - f0)
                with torch.cuda.stream(s_in):
                    sl["raw"][:g].copy_(h_raw[f0:f0 + g], non_blocking=True)
                    up = torch.cuda.Event()
                    up.record(s_in)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(up)
                    h_m[f0:f0 + g].copy_(sl["m"][:g], non_blocking=True)
                    h_c[f0:f0 + g].copy_(sl["c"][:g], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(s_out)
                if k >= 2:
                    s_in.wait_event(evs[k - 2])
                evs.append(done)
            s_out.synchronize()
            s_in.synchronize()

        for _ in range(3):
            copy_pass()
        t1 = time.perf_counter()
        copy_pass()
        one = max_over_ranks(time.perf_counter() - t1)  # the same number of passes on every rank
        passes = max(1, int(np.ceil(args.min_seconds / max(one, 1e-6) / max(args.steps, 1))))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps * passes):
            copy_pass()
        dt = time.perf_counter() - t0
        barrier()
        per_rank = gather_ranks(Be * rb * args.steps * passes / dt / 1e9)
        dt_max = max_over_ranks(dt)
        if rank == 0:
            emit({"mode": "copy-only", "n_gpus": world, "unit": "GB/s host->device per rank (pinned, 4-frame chunks, + records/counters back)",
                  "h2d_gbs_per_rank": per_rank, "h2d_gbs_total": sum(per_rank), "frames_per_s_equivalent": world * Be * args.steps * passes / dt_max,
                  "seconds": dt_max, "frame_bytes": rb, "frames_per_pass": Be, "passes": args.steps * passes,
                  "host_cpus_of_rank0": len(numa_cpus) if numa_cpus else "all", "config": workload_config(args, lp, world)})
        pin_raw.free(); pin_m.free(); pin_c.free()
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    d_raw = torch.empty((B, rb), dtype=torch.uint8, device=dev)
    h_frames = torch.from_numpy(frames)
    for i in range(B):
        d_raw[i].copy_(h_frames[i % n_distinct])
    d_flat = torch.empty((B, nf * 4), dtype=torch.uint8, device=dev)
    d_grad = torch.empty((B, nf), dtype=torch.float32, device=dev)
    d_circ = torch.empty((B, nf), dtype=torch.float32, device=dev)
    d_matches = torch.zeros((B, p.max_blobs * 22), dtype=torch.uint8, device=dev)
    d_counter = torch.zeros((B, 3), dtype=torch.int32, device=dev)
    nv12_stride = 2 * max(nf, nq)
    d_nv12 = torch.empty((B, nv12_stride), dtype=torch.uint8, device=dev) if args.config == 4 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if B * rb <= 126e6 else None
    torch.cuda.synchronize()
    import ctypes as C
    L = ctx.lib

    def one_pass():
        ctx.detect_batch_device(d_raw.data_ptr(), B, p, d_flat.data_ptr(), d_grad.data_ptr(), d_circ.data_ptr(), d_matches.data_ptr(), d_counter.data_ptr())
        if d_nv12 is not None:  # one view per frame, the four views over four contiguous quarters of the ring: four launches in all
            q = B // 4
            ctx._ck(L.vp_raw2nv12_batch_device(ctx.h, C.c_void_p(d_raw[0].data_ptr()), q, p.fmt, p.wq, p.hq, C.c_void_p(d_nv12[0].data_ptr()), nv12_stride, 0))
            ctx._ck(L.vp_rgba2nv12_batch_device(ctx.h, C.c_void_p(d_flat[q].data_ptr()), q, p.wf, p.hf, C.c_void_p(d_nv12[q].data_ptr()), nv12_stride))
            ctx._ck(L.vp_f2nv12_batch_device(ctx.h, C.c_void_p(d_grad[2 * q].data_ptr()), q, p.wf, p.hf, C.c_void_p(d_nv12[2 * q].data_ptr()), nv12_stride))
            ctx._ck(L.vp_f2nv12_batch_device(ctx.h, C.c_void_p(d_circ[3 * q].data_ptr()), B - 3 * q, p.wf, p.hf, C.c_void_p(d_nv12[3 * q].data_ptr()), nv12_stride))

    for _ in range(args.warmup):
        one_pass()
    ctx.sync()
    counters = d_counter.cpu().numpy()
    assert (counters[:, 0] > 0).all() and ctx.sat_fallbacks() == 0, "warm-up produced no blobs or left the exact SAT range"
    for i in range(n_distinct, B):  # identical frames must give identical counters
        assert (counters[i] == counters[i % n_distinct]).all()
    blobs_per_frame = float(np.minimum(counters[:, 0], p.max_blobs).mean())
    plan = ctx.last_plan()

    # ---- passes per step: the K timed steps last at least --min-seconds -----------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(4):
        one_pass()
    e1.record(stream)
    e1.synchronize()
    pass_ms = max_over_ranks(e0.elapsed_time(e1) / 4)
    passes = max(1, int(np.ceil(args.min_seconds * 1e3 / max(pass_ms, 1e-3) / max(args.steps, 1)))) if flush is None else 1

    def step():
        for _ in range(passes):
            one_pass()

    # ---- timed region: K steps, device-resident inputs -------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    if flush is None:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        e1.synchronize()
        my_ms = e0.elapsed_time(e1)
    else:  # a ring smaller than L2 (small --batch of the config-5 sweep): flush between steps, time every step on its own
        my_ms = 0.0
        for _ in range(args.steps):
            with torch.cuda.stream(stream):
                flush.fill_(1)
            e0.record(stream)
            step()
            e1.record(stream)
            e1.synchronize()
            my_ms += e0.elapsed_time(e1)
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    barrier()
    launches = ctx.launch_count() - launches0
    elapsed_ms = max_over_ranks(my_ms)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    frames_per_step = B * passes
    value = world * frames_per_step * args.steps / (elapsed_ms * 1e-3)

    # ---- the same pass with a CUDA event pair around every kernel; the library runs a profiled call on ONE stream, so a kernel's
    # duration is its own (the three-stream pipeline above overlaps them: the durations add up to more than its step) ----
    ctx.profiling(True)
    one_pass()
    ctx.sync()
    first = ctx.runtimes()
    n_prof = max(1, min(args.steps, 4000 // max(1, len(first))))
    for _ in range(n_prof):
        one_pass()
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
    frames_per_launch = B * n_prof / per_stage[top][1]  # frames of the profiled passes / launches of that kernel
    avg_ms = per_stage[top][0] / per_stage[top][1]
    achieved = abytes[top] * frames_per_launch / (avg_ms * 1e-3) / 1e9
    per_frame_traffic, traffic_src = traffic_from_capture(top)
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": per_frame_traffic * frames_per_launch if per_frame_traffic else None, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes[top] * frames_per_launch,
                "avg_launch_ms": avg_ms, "share_of_step": per_stage[top][0] / total_ms,
                "all_kernels": {k: {"frac": abytes[k] * (B * n_prof / v[1]) / (v[0] / v[1] * 1e-3) / 1e9 / peak_gbs, "us_per_frame": 1e3 * v[0] / (B * n_prof)}
                                for k, v in per_stage.items() if k in abytes and k != "peaks_emit"}}
    pipeline_gbs = abytes["frame"] * (value / world) / 1e9
    stage_ms = {k: {"ms_per_frame": v[0] / (B * n_prof), "share": v[0] / total_ms} for k, v in per_stage.items()}

    # ---- e2e: host frames in, blob lists out, through vp_detect_host ------------------------------
    def e2e_pass():
        ctx.detect_host_into(pin_raw.ptr.value, Be, p, pin_m.ptr.value, pin_c.ptr.value)

    for _ in range(3):
        e2e_pass()
    n_cmp = min(n_distinct, B, Be)
    assert (pin_c.array[:n_cmp] == counters[:n_cmp]).all(), "host path and device path disagree"
    t1 = time.perf_counter()
    e2e_pass()
    e2e_steps = max(3, min(args.steps, 10))
    e2e_passes = max(1, int(np.ceil(args.min_seconds / max(max_over_ranks(time.perf_counter() - t1), 1e-6) / e2e_steps)))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps * e2e_passes):
        e2e_pass()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_frames_per_step = Be * e2e_passes
    e2e_value = world * e2e_frames_per_step * e2e_steps / e2e_s

    # ---- single-frame latency through the host API (p50 / p99) ------------------------------------
    latency = None
    if not args.no_latency:
        lat = []
        replays0 = ctx.latency_graph_replays()
        for i in range(1020):
            t0 = time.perf_counter()
            ctx.detect_host_into(pin_raw.ptr.value + (i % Be) * rb, 1, p, pin_m.ptr.value, pin_c.ptr.value)
            lat.append(1e3 * (time.perf_counter() - t0))
        lat = np.sort(np.array(lat[20:]))
        latency = {"p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99)]), "frames": int(len(lat)),
                   "path": "vp_detect_host, 1 frame per call, pinned host in -> host blob list out; upload in 2 chunks of rows with the "
                           "reprojection of the first under the second, one download, replayed as a CUDA graph",
                   "graph_replays": int(ctx.latency_graph_replays() - replays0)}

    out = {
        "metric": metric_name(*args.size), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32/f32", "data": "synthetic",
        "config": workload_config(args, lp, world),
        "run": {"passes_per_step": passes, "frames_per_step_per_gpu": frames_per_step, "timed_seconds": elapsed_ms * 1e-3, "blobs_per_frame": blobs_per_frame,
                "plan": plan, "reprojection_tiles": ctx.tile_stats(p), "host_cpus_of_rank0": len(numa_cpus) if numa_cpus else "all",
                "note": "a step = passes_per_step passes of vp_detect_batch_device over the device-resident ring; every frame of every pass is counted"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_frames_per_step * rb, "d2h_bytes_per_step": e2e_frames_per_step * (p.max_blobs * 22 + 12),
                "frames_per_step": e2e_frames_per_step, "steps": e2e_steps, "timed_seconds": e2e_s,
                "h2d_gbs_per_gpu": e2e_value / world * rb / 1e9},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_pipeline": {"bound": "hbm", "achieved": pipeline_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": pipeline_gbs / peak_gbs,
                              "algorithmic_bytes_per_frame": abytes["frame"], "note": "per GPU; compulsory input + API outputs per frame x frames/s"
                              + ("; config 4 additionally writes 1.5 B per pixel of NV12 per frame, not counted here" if args.config == 4 else "")},
        "stage_ms": stage_ms,
        "latency": latency,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, lp, frames)
    if rank == 0:
        emit(out)
    pin_raw.free(); pin_m.free(); pin_c.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
