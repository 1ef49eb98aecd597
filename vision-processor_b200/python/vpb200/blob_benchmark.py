"""Ground-truth scoring of the detection images: the measurement half of the reference's blob_benchmark
(src/blob_benchmark.cpp:45-111 scoreBlob/scoreBot, :160-222 the per-frame loop and the summary lines).

The reference times raw2quad + rgba2blobCenter on every frame and then asks, for every blob the ground truth lists (balls,
and the five pattern blobs of every robot), where the arg-max circularity peak inside the blob's disc lies: position error
in field millimetres per blob colour, the summed peak circularity against the 99th percentile of the image ("worstblob /
percentile"), and a machine-readable ``[BlobMachine]`` line.  This module is that scoring, on the `blobCenter` image this
library produces -- the one check of the pipeline that does not depend on how oracle/clemu.h spells the sampler: a
half-texel convention error anywhere between raw2quad and satBlobCenter would move every peak off its blob.

CPU only (numpy) and independent of the CUDA library and of oracle/: it consumes images, wherever they came from.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as dc_field

import numpy as np

from .geometry import CENTER_BLOB_RADIUS, SIDE_BLOB_RADIUS, Perspective
from .synth import PATTERNS, PATTERN_POS, Scene

F32 = np.float32
ORANGE, YELLOW, BLUE, GREEN, PINK, BOT = range(6)  # enum BlobColor, blob_benchmark.cpp:24-31
COLOR_NAMES = ["ORANGE", "YELLOW", "BLUE", "GREEN", "PINK", "BOT"]
# patternAnglesb2b[5*i] (src/pattern.h): the angle of pattern blob i seen from the bot centre
PATTERN_ANGLES = [0.0] + [math.atan2(y, x) for (x, y) in PATTERN_POS[1:]]


@dataclass
class Accumulators:
    """The running sums of blob_benchmark.cpp:120-135."""
    frames: int = 0
    blob_amount: dict = dc_field(default_factory=lambda: {c: 0 for c in range(6)})
    offset_sum: dict = dc_field(default_factory=lambda: {c: np.zeros(2) for c in range(6)})
    error_sum: dict = dc_field(default_factory=lambda: {c: 0.0 for c in range(6)})
    error_sq_sum: dict = dc_field(default_factory=lambda: {c: 0.0 for c in range(6)})
    blob_score_sum: float = 0.0
    percentile_sum: float = 0.0
    max_error: float = 0.0
    missed: int = 0   # ground-truth blobs with no local peak inside their disc (the reference silently skips them)
    processing_time: float = 0.0


def field2flat(persp: Perspective, field_xyz, max_bot_height: float) -> np.ndarray:
    """blob_benchmark.cpp:40-42: field -> image -> back onto the plane at maxBotHeight -> flat pixel."""
    m = persp.model
    img = m.field2image(np.asarray(field_xyz, F32))
    return np.asarray(persp.field2flat(m.image2field(img, max_bot_height)[:2]), F32)


def score_blob(persp: Perspective, circ: np.ndarray, flat_xy, radius_px: float):
    """scoreBlob, blob_benchmark.cpp:45-84 with scoreMap == circMap (as the benchmark calls it, :167-179): the arg-max
    strict local peak inside the disc, refined by the parabola through its neighbours.  Returns (sub-pixel position,
    peak circularity) or None."""
    h, w = circ.shape
    fx, fy = float(flat_xy[0]), float(flat_xy[1])
    best, best_pos = -math.inf, None
    for y in range(max(0, int(math.floor(fy - radius_px))), min(h, int(math.ceil(fy + radius_px)))):
        xr = math.sqrt(max(radius_px * radius_px - (y - fy) * (y - fy), 0.0))
        for x in range(max(0, int(math.floor(fx - xr))), min(w, int(math.ceil(fx + xr)))):
            s = circ[y, x]
            if s > best:
                c = s
                nx, px = circ[y, max(0, x - 1)], circ[y, min(w - 1, x + 1)]
                ny, py = circ[max(0, y - 1), x], circ[min(h - 1, y + 1), x]
                if c > nx and c > px and c > ny and c > py:
                    xdiv, ydiv = nx - 2 * c + px, ny - 2 * c + py
                    best_pos = (x + (0.5 * (nx - px) / xdiv if xdiv != 0 else 0.0), y + (0.5 * (ny - py) / ydiv if ydiv != 0 else 0.0))
                    best = s
    return None if best_pos is None else (np.asarray(best_pos, F32), float(best))


def _account(acc: Accumulators, persp: Perspective, circ, field_xyz, radius_mm, color, max_bot_height):
    flat = field2flat(persp, field_xyz, max_bot_height)
    hit = score_blob(persp, circ, flat, radius_mm / persp.field_scale)
    if hit is None:
        acc.missed += 1
        return np.zeros(2), 0.0
    pos, score = hit
    offset = np.asarray(persp.flat2field(pos), np.float64) - np.asarray(persp.flat2field(flat), np.float64)
    n = float(np.hypot(*offset))
    acc.blob_amount[color] += 1
    acc.offset_sum[color] += offset
    acc.error_sum[color] += n
    acc.error_sq_sum[color] += n * n
    acc.max_error = max(acc.max_error, n)
    return offset, score


def score_frame(acc: Accumulators, persp: Perspective, circ: np.ndarray, scene: Scene, max_bot_height: float = 180.0, processing_time: float = 0.0):
    """One iteration of the frame loop, blob_benchmark.cpp:160-194, for the ground truth of `scene`."""
    blob_score = 0.0
    for b in scene.balls:  # balls at z = 30 like ssl-vision (:166)
        _, s = _account(acc, persp, circ, (b.x, b.y, 30.0), scene.field.ball_radius, ORANGE, max_bot_height)
        blob_score += s
    for r in scene.robots:  # scoreBot, :86-111
        bot_color = YELLOW if r.team == "yellow" else BLUE
        pat = PATTERNS[r.robot_id]
        bot_offset = np.zeros(2)
        for i in range(5):
            ori = r.orientation + PATTERN_ANGLES[i]
            dist = math.hypot(*PATTERN_POS[i])
            color = bot_color if i == 0 else (GREEN if pat & (8 >> i) else PINK)
            off, s = _account(acc, persp, circ, (r.x + dist * math.cos(ori), r.y + dist * math.sin(ori), r.height),
                              CENTER_BLOB_RADIUS if i == 0 else SIDE_BLOB_RADIUS, color, max_bot_height)
            blob_score += s
            bot_offset += off / 5
        acc.blob_amount[BOT] += 1
        acc.offset_sum[BOT] += bot_offset
        acc.error_sum[BOT] += float(np.hypot(*bot_offset))
        acc.error_sq_sum[BOT] += float(bot_offset @ bot_offset)
    flat = np.sort(circ.reshape(-1))  # nth_element at 99 % of the image, :190-192 (dense pitch)
    acc.percentile_sum += float(flat[int(circ.size * 0.99)])
    acc.blob_score_sum += blob_score
    acc.processing_time += processing_time
    acc.frames += 1


def summary(acc: Accumulators, persp: Perspective) -> dict:
    """The closing arithmetic of blob_benchmark.cpp:196-222 and its two output lines."""
    total_error = sum(acc.error_sum[c] for c in range(5))
    total_sq = sum(acc.error_sq_sum[c] for c in range(5))
    total_blobs = sum(acc.blob_amount[c] for c in range(5))
    score = acc.blob_score_sum / max(total_blobs, 1)
    stddev = math.sqrt(max(total_blobs * total_sq - total_error * total_error, 0.0)) / max(total_blobs, 1)
    ppr = score / (abs(score) + abs(acc.percentile_sum)) if (score or acc.percentile_sum) else 0.0
    machine = "[BlobMachine] " + " ".join(str(v) for v in [
        acc.frames, total_blobs, total_error, total_sq, score, acc.percentile_sum,
        acc.blob_amount[ORANGE], acc.error_sum[ORANGE], acc.error_sq_sum[ORANGE],
        acc.blob_amount[BOT], acc.error_sum[BOT], acc.error_sq_sum[BOT], total_blobs * persp.field_scale, acc.processing_time])
    return {
        "frames": acc.frames, "blobs": total_blobs, "missed": acc.missed,
        "mean_error_mm": total_error / max(total_blobs, 1), "stddev_mm": stddev, "max_error_mm": acc.max_error,
        "mean_error_flat_px": total_error / max(total_blobs, 1) / persp.field_scale,
        "worstblob_percentile": ppr,
        "per_color": {COLOR_NAMES[c]: {"n": acc.blob_amount[c], "mean_error_mm": acc.error_sum[c] / max(acc.blob_amount[c], 1),
                                       "systematic_offset_mm": (acc.offset_sum[c] / max(acc.blob_amount[c], 1)).tolist()} for c in range(6)},
        "lines": [f"[Blob benchmark] Total error: {total_error / max(total_blobs, 1)}±{stddev} worstblob/percentile: {ppr}", machine],
    }
