"""Deterministic synthetic SSL camera frames (the reference has no renderer).

Renders what a pinhole+k2 camera (``geometry.CameraModel``) sees of a Division-A-like field --
carpet, white lines, robots with the butterfly pattern of src/pattern.h:19-56, orange balls --
as a raw Bayer (RGGB/GRBG) or BGR frame, plus ground truth in the schema of
src/GroundTruth.cpp:23-78.  Everything is seeded; the same (scene, camera, seed) gives the same bytes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as dc_field

import numpy as np

from .geometry import CameraModel, FieldSize, CENTER_BLOB_RADIUS, SIDE_BLOB_RADIUS

F32 = np.float32

FMT_RGGB, FMT_GRBG, FMT_BGR = 0, 1, 2

# src/pattern.h:19-36  (1 = green, 0 = pink; MSB first, increasing angle from the bot orientation)
PATTERNS = [0b0100, 0b1100, 0b1101, 0b0101, 0b0010, 0b1010, 0b1011, 0b0011,
            0b1111, 0b0000, 0b0110, 0b1001, 0b1110, 0b1000, 0b0111, 0b0001]
# src/pattern.h:47-53
PATTERN_POS = [(0.0, 0.0), (35.0, 54.772), (-54.772, 35.0), (-54.772, -35.0), (35.0, -54.772)]

# sensor RGB of the scene materials
CARPET = (52.0, 104.0, 60.0)
LINE = (225.0, 225.0, 225.0)
BOT_TOP = (24.0, 24.0, 24.0)
ORANGE = (250.0, 120.0, 24.0)
YELLOW = (235.0, 225.0, 40.0)
BLUE = (40.0, 90.0, 235.0)
GREEN = (60.0, 225.0, 95.0)
PINK = (245.0, 70.0, 170.0)


@dataclass
class Robot:
    team: str          # "yellow" | "blue"
    robot_id: int
    x: float
    y: float
    orientation: float
    height: float = 145.0


@dataclass
class Ball:
    x: float
    y: float
    z: float = 21.5


@dataclass
class Scene:
    field: FieldSize = dc_field(default_factory=FieldSize)
    robots: list = dc_field(default_factory=list)
    balls: list = dc_field(default_factory=list)

    def ground_truth(self) -> dict:
        """One frame of the YAML schema parsed by src/GroundTruth.cpp:23-78."""
        out = {"balls": [{"x": float(b.x), "y": float(b.y)} for b in self.balls], "robots_yellow": [], "robots_blue": []}
        for r in self.robots:
            out["robots_" + r.team].append({"robot_id": int(r.robot_id), "x": float(r.x), "y": float(r.y),
                                            "orientation": float(r.orientation), "height": float(r.height)})
        return out

    def detection_frame(self, model: CameraModel, camera_id: int = 0, frame_number: int = 1, t_capture: float = 0.0) -> dict:
        """One SSL_DetectionFrame with every field src/GroundTruth.cpp:23-78 reads (the decoders call .as<> on confidence,
        x, y, pixel_x, pixel_y without a default, so those must be present): pixel positions through the camera model."""
        def px(x, y, z):
            p = model.field2image(np.array([x, y, z], F32))
            return float(p[0]), float(p[1])
        frame = {"camera_id": int(camera_id), "frame_number": int(frame_number), "t_capture": float(t_capture), "t_sent": float(t_capture),
                 "balls": [], "robots_blue": [], "robots_yellow": []}
        for b in self.balls:
            u, v = px(b.x, b.y, b.z)
            frame["balls"].append({"confidence": 1.0, "x": float(b.x), "y": float(b.y), "z": float(b.z), "pixel_x": u, "pixel_y": v})
        for r in self.robots:
            u, v = px(r.x, r.y, r.height)
            frame["robots_" + r.team].append({"confidence": 1.0, "robot_id": int(r.robot_id), "x": float(r.x), "y": float(r.y),
                                              "orientation": float(r.orientation), "pixel_x": u, "pixel_y": v, "height": float(r.height)})
        return frame

    def blobs(self) -> list:
        """(x, y, z, radius_mm, rgb) of every coloured disc, robots first (scoreBot, blob_benchmark.cpp:86-111)."""
        out = []
        for r in self.robots:
            pat = PATTERNS[r.robot_id]
            c, s = math.cos(r.orientation), math.sin(r.orientation)
            for i, (px, py) in enumerate(PATTERN_POS):
                col = (YELLOW if r.team == "yellow" else BLUE) if i == 0 else (GREEN if pat & (8 >> i) else PINK)
                out.append((r.x + c * px - s * py, r.y + s * px + c * py, r.height,
                            CENTER_BLOB_RADIUS if i == 0 else SIDE_BLOB_RADIUS, col))
        for b in self.balls:
            out.append((b.x, b.y, b.z, self.field.ball_radius, ORANGE))
        return out


def random_scene(extent, n_robots: int, n_balls: int, seed: int, fieldsize: FieldSize | None = None) -> Scene:
    """Robots/balls placed on a jittered grid inside extent=(xmin,xmax,ymin,ymax) so they never overlap."""
    rng = np.random.default_rng(seed)
    fs = fieldsize or FieldSize()
    xmin, xmax, ymin, ymax = extent
    margin = 130.0
    n = n_robots + n_balls
    cols = max(1, int(math.ceil(math.sqrt(n * (xmax - xmin) / max(ymax - ymin, 1.0)))))
    rows = max(1, int(math.ceil(n / cols)))
    cw, chh = (xmax - xmin - 2 * margin) / cols, (ymax - ymin - 2 * margin) / rows
    cells = [(i, j) for j in range(rows) for i in range(cols)]
    order = rng.permutation(len(cells))
    sc = Scene(field=fs)
    for k in range(min(n, len(cells))):
        i, j = cells[order[k]]
        jx = (rng.random() - 0.5) * max(cw - 2 * 100.0, 0.0)
        jy = (rng.random() - 0.5) * max(chh - 2 * 100.0, 0.0)
        x = xmin + margin + (i + 0.5) * cw + jx
        y = ymin + margin + (j + 0.5) * chh + jy
        if k < n_robots:
            sc.robots.append(Robot("yellow" if k % 2 == 0 else "blue", int(rng.integers(0, 16)), float(x), float(y),
                                   float(rng.uniform(-math.pi, math.pi)), float(rng.choice([140.0, 145.0, 150.0]))))
        else:
            sc.balls.append(Ball(float(x), float(y)))
    return sc


def _seg_dist(px, py, x0, y0, x1, y1):
    dx, dy = x1 - x0, y1 - y0
    l2 = dx * dx + dy * dy
    t = np.clip(((px - x0) * dx + (py - y0) * dy) / l2, 0.0, 1.0)
    return np.hypot(px - (x0 + t * dx), py - (y0 + t * dy))


def _blend(img, cov, rgb):
    for c in range(3):
        img[..., c] += cov * (F32(rgb[c]) - img[..., c])


def render_rgb(scene: Scene, model: CameraModel, sensor_w: int, sensor_h: int) -> np.ndarray:
    """Noise-free sensor-resolution RGB (float32, HxWx3).  Sensor pixel (X,Y) sits at quad coordinate
    ((X+.5)/2, (Y+.5)/2): quad (i,j) holds R,G1,G2,B at (+.25,+.25),(+.75,+.25),(+.25,+.75),(+.75,+.75)
    -- the geometry the +-0.25 taps of resampling.cl:65-70 assume."""
    bayer_div = 2.0 if sensor_w == 2 * model.size[0] else 1.0
    ys, xs = np.mgrid[0:sensor_h, 0:sensor_w].astype(F32)
    q = np.stack([(xs + F32(0.5)) / F32(bayer_div), (ys + F32(0.5)) / F32(bayer_div)], -1)
    img = np.empty((sensor_h, sensor_w, 3), F32)
    img[...] = np.asarray(CARPET, F32)

    ground = model.image2field(q, 0.0)
    gx, gy = ground[..., 0], ground[..., 1]
    # mm per sensor pixel (for one-pixel-wide anti-aliased edges)
    px_mm = np.maximum(np.hypot(np.gradient(gx, axis=1), np.gradient(gy, axis=1)), F32(1e-3))

    f = scene.field
    hl, hw, t = f.field_length / 2, f.field_width / 2, f.line_thickness / 2
    segs = [(-hl, -hw, hl, -hw), (-hl, hw, hl, hw), (-hl, -hw, -hl, hw), (hl, -hw, hl, hw),
            (0, -hw, 0, hw), (-hl, 0, hl, 0)]
    pw, pd = f.penalty_area_width / 2, f.penalty_area_depth
    for sgn in (-1, 1):
        segs += [(sgn * hl, -pw, sgn * (hl - pd), -pw), (sgn * hl, pw, sgn * (hl - pd), pw),
                 (sgn * (hl - pd), -pw, sgn * (hl - pd), pw)]
    d = np.full(gx.shape, np.inf, F32)
    for s in segs:
        d = np.minimum(d, _seg_dist(gx, gy, *[F32(v) for v in s]).astype(F32))
    d = np.minimum(d, np.abs(np.hypot(gx, gy) - F32(f.center_circle_radius)))
    _blend(img, np.clip((F32(t) - d) / px_mm + F32(0.5), 0, 1), LINE)

    def disc(cx, cy, cz, radius, rgb):
        c_img = model.field2image(np.array([cx, cy, cz], F32)) * F32(bayer_div)
        r_px = radius * model.focal_length / max(model.pos[2] - cz, 1.0) * bayer_div * 1.6 + 4
        x0, x1 = int(max(0, math.floor(c_img[0] - r_px))), int(min(sensor_w, math.ceil(c_img[0] + r_px)))
        y0, y1 = int(max(0, math.floor(c_img[1] - r_px))), int(min(sensor_h, math.ceil(c_img[1] + r_px)))
        if x1 <= x0 or y1 <= y0:
            return
        p = model.image2field(q[y0:y1, x0:x1], cz)
        dd = np.hypot(p[..., 0] - F32(cx), p[..., 1] - F32(cy))
        _blend(img[y0:y1, x0:x1], np.clip((F32(radius) - dd) / px_mm[y0:y1, x0:x1] + F32(0.5), 0, 1), rgb)

    for r in scene.robots:
        disc(r.x, r.y, r.height, f.max_robot_radius, BOT_TOP)
    for (x, y, z, rad, rgb) in scene.blobs():
        disc(x, y, z, rad, rgb)
    return img


def mosaic(rgb: np.ndarray, fmt: int) -> np.ndarray:
    """RGB (HxWx3 uint8) -> raw frame bytes.  RGGB: R at even row/even col (raw2quad.cl:34-37)."""
    if fmt == FMT_BGR:
        return np.ascontiguousarray(rgb[..., ::-1])
    h, w = rgb.shape[:2]
    raw = np.empty((h, w), np.uint8)
    if fmt == FMT_RGGB:
        raw[0::2, 0::2] = rgb[0::2, 0::2, 0]
        raw[0::2, 1::2] = rgb[0::2, 1::2, 1]
        raw[1::2, 0::2] = rgb[1::2, 0::2, 1]
        raw[1::2, 1::2] = rgb[1::2, 1::2, 2]
    else:  # GRBG
        raw[0::2, 0::2] = rgb[0::2, 0::2, 1]
        raw[0::2, 1::2] = rgb[0::2, 1::2, 0]
        raw[1::2, 0::2] = rgb[1::2, 0::2, 2]
        raw[1::2, 1::2] = rgb[1::2, 1::2, 1]
    return raw


def add_noise(clean: np.ndarray, seed: int, sigma: float = 2.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    noisy = clean + rng.normal(0.0, sigma, clean.shape).astype(F32)
    return np.clip(np.rint(noisy), 0, 255).astype(np.uint8)


def render_raw(scene: Scene, model: CameraModel, sensor_w: int, sensor_h: int, fmt: int = FMT_RGGB,
               seed: int = 0, sigma: float = 2.0, clean_rgb: np.ndarray | None = None) -> np.ndarray:
    """One raw frame.  seed = 1000*cam + frame by convention.  Pass clean_rgb to re-noise a cached scene."""
    if clean_rgb is None:
        clean_rgb = render_rgb(scene, model, sensor_w, sensor_h)
    return mosaic(add_noise(clean_rgb, seed, sigma), fmt)


def noise_frame(sensor_w: int, sensor_h: int, seed: int, fmt: int = FMT_RGGB) -> np.ndarray:
    """Stress frame: uniform random bytes (many rejected peaks, exercises counter[2])."""
    rng = np.random.default_rng(seed)
    shape = (sensor_h, sensor_w, 3) if fmt == FMT_BGR else (sensor_h, sensor_w)
    return rng.integers(0, 256, shape, dtype=np.uint8)


def write_ground_truth_yaml(path: str, frames: list) -> None:
    """`frames` (Scene.detection_frame dicts) as the YAML sequence parseGroundTruth loads (src/GroundTruth.cpp:81-83:
    YAML::LoadFile(source).as<std::vector<SSL_DetectionFrame>>()).  Plain block style, no anchors."""
    import yaml
    with open(path, "w") as f:
        yaml.safe_dump(frames, f, default_flow_style=False, sort_keys=False)


def parse_ground_truth_yaml(path: str) -> list:
    """What src/GroundTruth.cpp:23-78 extracts, with its required/optional split (a missing required key raises like
    yaml-cpp's .as<> would); used to check that what write_ground_truth_yaml emits is what the reference reads."""
    import yaml
    out = []
    for node in yaml.safe_load(open(path)):
        fr = {k: node[k] for k in ("camera_id", "frame_number", "t_capture", "t_sent")}
        if "t_capture_camera" in node:
            fr["t_capture_camera"] = node["t_capture_camera"]
        fr["balls"] = [dict({k: b[k] for k in ("confidence", "x", "y", "pixel_x", "pixel_y")}, **{k: b[k] for k in ("area", "z") if k in b}) for b in node["balls"]]
        for team in ("robots_blue", "robots_yellow"):
            fr[team] = [dict({k: r[k] for k in ("confidence", "x", "y", "pixel_x", "pixel_y")}, **{k: r[k] for k in ("robot_id", "orientation", "height") if k in r})
                        for r in node[team]]
        out.append(fr)
    return out
