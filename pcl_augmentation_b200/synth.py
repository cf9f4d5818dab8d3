"""Seeded synthetic inputs in the reference's on-disk array layouts.

Everything here is plain numpy and deterministic in ``seed``.  The generators make
* HDL-64 / OS1-128 shaped scans (``N x 4`` float32 + ``uint32`` semantic labels, ring-major like KITTI
  ``velodyne/*.bin`` read by reference ``object_detection/Real3DAug/tools/datasets.py:56-71``),
* rich maps in the two formats the placement search reads
  (OD ``{map uint8, min_x, min_y}``, reference ``object_detection/rich_map/single_drivable_area_map.py:161``;
  semseg ``{map float64 in {0,1,2,3}, move 4x1 int}``, reference
  ``semantic_segmentation/rich_map/drivable_area_map.py:205-206``),
* cut-object samples ``{pcl: M x 5 float64, anno: 0-d str}`` (reference
  ``object_detection/cut_object/object_cut_out.py:164-168``, ``semantic_segmentation/cut_object/cut_out.py:156-157``),
* scene box label lines, and
* the per-scan "RNG-drawn candidates" (class counts of ``generate_seed`` and the shuffled sample order of
  ``random.shuffle``, reference ``object_detection/Real3DAug/insertion.py:171-187, 400``) as explicit tables so the
  oracle and the CUDA path consume identical draws.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SENSOR_HEIGHT = 1.73
ROAD_HALF_WIDTH = 4.0
SIDEWALK_WIDTH = 2.0
CROSS_X0, CROSS_X1 = 16.0, 24.0

# KITTI-style (h, w, l) with l along the heading
OD_DIMS = {"Pedestrian": (1.8, 0.6, 0.8), "Cyclist": (1.7, 0.6, 1.8), "Car": (1.5, 1.6, 3.9)}
SS_DIMS = {11: (1.1, 0.6, 1.7), 15: (1.4, 0.8, 2.0), 18: (3.0, 2.5, 7.0), 30: (1.8, 0.6, 0.8),
           31: (1.7, 0.6, 1.8), 32: (1.6, 0.8, 2.0), 253: (1.7, 0.6, 1.8), 255: (1.6, 0.8, 2.0)}


@dataclass
class ScanShape:
    beams: int = 64
    az_steps: int = 1875
    el_top_deg: float = 2.0
    el_bot_deg: float = -24.8


KITTI_SHAPE = ScanShape(64, 1875, 2.0, -24.8)            # 120 000 pts   (config 1 / 3)
SEMKITTI_SHAPE = ScanShape(64, 1953, 2.0, -24.8)         # 124 992 pts   (config 2 / 5)
OS128_SHAPE = ScanShape(128, 2048, 22.5, -22.5)          # 262 144 pts   (config 4)
SMALL_SHAPE = ScanShape(16, 450, 2.0, -24.8)             # 7 200 pts     (parity tests vs the real reference)


def ground_label(x, y):
    """Semantic label of a ground point at lidar-frame (x, y): 40 road, 48 sidewalk, 44 parking, 72 terrain."""
    x = np.asarray(x)
    y = np.asarray(y)
    road = (np.abs(y) < ROAD_HALF_WIDTH) | ((x > CROSS_X0) & (x < CROSS_X1))
    side = (~road) & ((np.abs(y) < ROAD_HALF_WIDTH + SIDEWALK_WIDTH)
                      | ((x > CROSS_X0 - SIDEWALK_WIDTH) & (x < CROSS_X1 + SIDEWALK_WIDTH)))
    park = (~road) & (~side) & (x < -10) & (x > -22) & (y > 6) & (y < 12)
    out = np.full(x.shape, 72, dtype=np.uint32)
    out[park] = 44
    out[side] = 48
    out[road] = 40
    return out


def _ray_dirs(shape: ScanShape, rng, az_jitter=0.25):
    el = np.deg2rad(np.linspace(shape.el_top_deg, shape.el_bot_deg, shape.beams))
    az = 2 * np.pi * np.arange(shape.az_steps) / shape.az_steps
    el = np.repeat(el, shape.az_steps)
    az = np.tile(az, shape.beams)
    az = az + rng.uniform(-az_jitter, az_jitter, az.shape) * (2 * np.pi / shape.az_steps)
    el = el + rng.normal(0.0, 2e-4, el.shape)
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    return d, az


def _ellipsoid_hits(dirs, centre, yaw, semi):
    """Ray (from the origin, unit ``dirs``) vs. yaw-rotated ellipsoid; returns t (inf where missed)."""
    c, s = math.cos(yaw), math.sin(yaw)
    rot = np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])     # world -> local
    inv = 1.0 / np.asarray(semi, dtype=np.float64)
    dl = (dirs @ rot.T) * inv
    ol = (rot @ (-np.asarray(centre, dtype=np.float64))) * inv
    a = np.einsum("ij,ij->i", dl, dl)
    b = 2.0 * (dl @ ol)
    cc = float(ol @ ol) - 1.0
    disc = b * b - 4 * a * cc
    t = np.full(len(dirs), np.inf)
    ok = disc > 0
    sq = np.sqrt(disc[ok])
    t0 = (-b[ok] - sq) / (2 * a[ok])
    t0[t0 <= 0] = np.inf
    t[ok] = t0
    return t


def _wall_ranges(az, rng):
    """Piecewise-constant wall range per azimuth, kept clear of the road + sidewalk corridor."""
    edges = [0.0]
    while edges[-1] < 2 * np.pi:
        edges.append(edges[-1] + np.deg2rad(rng.uniform(2.0, 15.0)))
    edges = np.array(edges)
    seg_r = rng.uniform(8.0, 48.0, len(edges))
    seg_lab = rng.choice(np.array([50, 51, 70, 71, 80], dtype=np.uint32), len(edges))
    seg = np.clip(np.searchsorted(edges, np.mod(az, 2 * np.pi), side="right") - 1, 0, len(edges) - 1)
    w = seg_r[seg]
    clear = (ROAD_HALF_WIDTH + SIDEWALK_WIDTH + 1.0) / np.maximum(np.abs(np.sin(az)), 1e-3)
    w = np.minimum(np.maximum(w, clear), 70.0)
    return w, seg_lab[seg]


def make_scene_cars(seed, count):
    """``count`` car boxes parked on the road corridor: list of (cx, cy, z_bottom, yaw_heading, h, w, l)."""
    rng = np.random.default_rng([seed, 7])
    cars = []
    for i in range(count):
        x = rng.uniform(-45, 45)
        if abs(x) < 6:
            x += 12 * np.sign(x if x != 0 else 1.0)
        y = rng.choice([-2.6, 2.6]) + rng.uniform(-0.3, 0.3)
        yaw = rng.choice([0.0, np.pi]) + rng.uniform(-0.05, 0.05)
        h, w, l = OD_DIMS["Car"]
        cars.append((float(x), float(y), -SENSOR_HEIGHT, float(yaw), h, w, l))
    return cars


def make_scan(seed, shape: ScanShape = KITTI_SHAPE, cars=()):
    """One synthetic spinning-LiDAR scan.  Returns (pcl float32 N x 4, labels uint32 N), ring-major order."""
    rng = np.random.default_rng([seed, 1])
    dirs, az = _ray_dirs(shape, rng)
    n = len(dirs)
    with np.errstate(divide="ignore"):
        t_ground = np.where(dirs[:, 2] < -1e-6, SENSOR_HEIGHT / -dirs[:, 2], np.inf)
    wall_r, wall_lab = _wall_ranges(az, rng)
    t_wall = wall_r / np.maximum(np.hypot(dirs[:, 0], dirs[:, 1]), 1e-9)
    t = np.minimum(t_ground, t_wall)
    labels = np.where(t_ground <= t_wall, 0, wall_lab).astype(np.uint32)
    for (cx, cy, zb, yaw, h, w, l) in cars:
        tc = _ellipsoid_hits(dirs, (cx, cy, zb + h / 2), yaw, (l / 2, w / 2, h / 2))
        hit = tc < t
        t = np.where(hit, tc, t)
        labels[hit] = 10
    t = t * (1.0 + rng.normal(0.0, 0.002, n))
    pts = dirs * t[:, None]
    g = labels == 0
    labels[g] = ground_label(pts[g, 0], pts[g, 1])
    pts[g, 2] = -SENSOR_HEIGHT + rng.normal(0.0, 0.01, int(g.sum()))
    inten = rng.uniform(0.0, 1.0, n)
    pcl = np.concatenate([pts, inten[:, None]], axis=1).astype(np.float32)
    return pcl, labels


# ----------------------------------------------------------------------------------------------- maps
def make_od_maps(extent=64):
    """(road_map, sidewalk_map) dicts ``{map: uint8 X x Y, min_x: int, min_y: int}``, 1 m cells."""
    xs = np.arange(-extent, extent) + 0.5
    gx, gy = np.meshgrid(xs, xs, indexing="ij")
    lab = ground_label(gx, gy)
    # a cell counts as road / sidewalk only if its four corners agree (keeps objects off the kerb line)
    def full(val):
        m = lab == val
        for dx in (-0.49, 0.49):
            for dy in (-0.49, 0.49):
                m &= ground_label(gx + dx, gy + dy) == val
        return m.astype(np.uint8)
    road = {"map": full(40), "min_x": np.int64(-extent), "min_y": np.int64(-extent)}
    side = {"map": full(48), "min_x": np.int64(-extent), "min_y": np.int64(-extent)}
    return road, side


def make_pose(seed, tilt=True):
    """4x4 lidar->world matrix in the role of ``SemanticKITTI.create_transform_matrix`` (reference
    ``semantic_segmentation/Real3DAug/tools/datasets.py:65-70``): yaw + small pitch/roll + translation."""
    rng = np.random.default_rng([seed, 2])
    yaw = rng.uniform(-np.pi, np.pi)
    pitch, roll = (rng.uniform(-0.02, 0.02, 2) if tilt else (0.0, 0.0))
    cy, sy, cp, sp, cr, sr = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch), math.cos(roll), math.sin(roll)
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    t = np.eye(4)
    t[:3, :3] = rz @ ry @ rx
    t[:3, 3] = [rng.uniform(-200, 200), rng.uniform(-200, 200), rng.uniform(-1, 1)]
    return t


def make_ss_map(pose, extent=96, margin=40):
    """Semseg rich map ``{map: float64 X x Y in {0,1,2,3}, move: int 4x1}`` covering the scan at ``pose``."""
    xs = np.arange(-extent, extent, 0.25) + 0.125
    gx, gy = np.meshgrid(xs, xs, indexing="ij")
    lab = ground_label(gx, gy).ravel()
    pts = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, -SENSOR_HEIGHT), np.ones(gx.size)])
    w = pose @ pts
    minx = int(np.floor(w[0].min())) - margin
    miny = int(np.floor(w[1].min())) - margin
    sx = int(np.ceil(w[0].max())) - minx + margin
    sy = int(np.ceil(w[1].max())) - miny + margin
    ix = (w[0] - minx).astype(np.int64)
    iy = (w[1] - miny).astype(np.int64)
    # a cell is kept only if every sample falling in it has the same surface (no kerb-straddling cells)
    val = np.select([lab == 40, lab == 48, lab == 44], [1, 2, 3], 0)
    flat = ix * sy + iy
    lo = np.full(sx * sy, 9, dtype=np.int64)
    hi = np.full(sx * sy, -1, dtype=np.int64)
    np.minimum.at(lo, flat, val)
    np.maximum.at(hi, flat, val)
    grid = np.where((lo == hi), hi, 0).astype(np.float64).reshape(sx, sy)
    grid[grid < 0] = 0
    move = np.array([[minx], [miny], [0], [1]], dtype=np.int64)
    return {"map": grid, "move": move}


# ------------------------------------------------------------------------------------------ cut objects
def _wrap_pi(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def od_label_line(cls, centre_bottom, yaw_heading, dims, truncated=0.0):
    """KITTI ``label_2`` line (15 fields, camera frame) whose ``read_label_line`` parse (reference
    ``object_detection/Real3DAug/tools/find_spot.py:175-224``) gives back the lidar-frame box."""
    h, w, l = dims
    bx, by, bz = centre_bottom
    cam = (-by, -(bz + 0.08), bx - 0.27)
    ry = _wrap_pi(-(yaw_heading - np.pi / 2))
    alpha = _wrap_pi(-math.atan2(cam[0], cam[2]) + ry)
    return (f"{cls} {truncated:.2f} 0 {alpha:.2f} 100.00 120.00 150.00 220.00 {h:.2f} {w:.2f} {l:.2f} "
            f"{cam[0]:.2f} {cam[1]:.2f} {cam[2]:.2f} {ry:.2f}")


def ss_label_line(cls, centre_bottom, yaw_heading, dims):
    """Semseg bbox line ``cls x y z h w l yaw`` parsed by reference
    ``semantic_segmentation/Real3DAug/tools/find_spot.py:155-189``."""
    h, w, l = dims
    bx, by, bz = centre_bottom
    return f"{cls} {bx:.4f} {by:.4f} {bz:.4f} {h + 0.1:.3f} {w + 0.1:.3f} {l + 0.1:.3f} {_wrap_pi(yaw_heading):.4f}"


def make_cut_object(seed, cls, semseg=False, shape: ScanShape = KITTI_SHAPE, rng_range=(5.0, 35.0)):
    """One cut-out object as the ``.npz`` payload ``{pcl: M x 5 float64, anno: 0-d str}``."""
    rng = np.random.default_rng([seed, 3])
    dims = (SS_DIMS if semseg else OD_DIMS)[cls]
    h, w, l = dims
    for _ in range(64):
        d = rng.uniform(*rng_range)
        phi = rng.uniform(0, 2 * np.pi)
        yaw = rng.uniform(-np.pi, np.pi)
        zb = -SENSOR_HEIGHT + rng.uniform(-0.03, 0.03)
        cx, cy = d * math.cos(phi), d * math.sin(phi)
        daz = 2 * np.pi / shape.az_steps
        half = math.atan2(0.6 * max(l, w), d) + 2 * daz
        az = np.arange(phi - half, phi + half, daz) + rng.uniform(0, daz)
        el = np.deg2rad(np.linspace(shape.el_top_deg, shape.el_bot_deg, shape.beams))
        el, az = np.meshgrid(el, az, indexing="ij")
        el = el.ravel() + rng.normal(0, 2e-4, el.size)
        az = az.ravel() + rng.uniform(-0.25, 0.25, az.size) * daz
        dirs = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
        t = _ellipsoid_hits(dirs, (cx, cy, zb + h / 2), yaw, (0.47 * l, 0.47 * w, 0.485 * h))
        ok = np.isfinite(t)
        if ok.sum() >= 5:
            break
    t = t[ok] * (1.0 + rng.normal(0, 0.001, int(ok.sum())))
    pts = (dirs[ok] * t[:, None]).astype(np.float32).astype(np.float64)
    inten = rng.uniform(0, 1, len(pts)).astype(np.float32).astype(np.float64)
    lab = np.full(len(pts), float(cls) if semseg else 1.0)
    pcl = np.concatenate([pts, inten[:, None], lab[:, None]], axis=1)
    line = (ss_label_line(cls, (cx, cy, zb), yaw, dims) if semseg
            else od_label_line(cls, (cx, cy, zb), yaw, dims))
    return {"pcl": pcl, "anno": np.array(line)}


def make_object_db(seed, classes, n_per_class, semseg=False, shape: ScanShape = KITTI_SHAPE, rng_range=(5.0, 35.0)):
    """``{class: [(name, sample), ...]}`` sorted by name — the role of ``glob(sample_path/<class>/*.npz)``."""
    db = {}
    for ci, cls in enumerate(classes):
        items = []
        for j in range(n_per_class):
            s = make_cut_object(seed * 100003 + ci * 1009 + j, cls, semseg, shape, rng_range)
            items.append((f"{cls}_{j:05d}", s))
        db[cls] = items
    return db


def make_scene_box_lines(cars, semseg=False):
    out = []
    for (cx, cy, zb, yaw, h, w, l) in cars:
        if semseg:
            out.append(ss_label_line(10, (cx, cy, zb), yaw, (h, w, l)))
        else:
            out.append(od_label_line("Car", (cx, cy, zb), yaw, (h, w, l)))
    return out


# ------------------------------------------------------------------------------------------- schedules
@dataclass
class Schedule:
    """Pre-drawn randomness of one scan: ``counts[c]`` = objects to insert per class (``generate_seed``),
    ``perms[e][c]`` = first ``MAX_NUM_TRIES`` indices of the e-th ``random.shuffle`` of class c's sorted list."""
    counts: np.ndarray
    perms: np.ndarray            # int32 [events, classes, tries]
    meta: dict = field(default_factory=dict)


def make_schedule(seed, n_classes, number_of_object, list_lens, tries=100, random_counts=True, fixed_counts=None):
    rng = np.random.default_rng([seed, 4])
    if random_counts:
        counts = np.zeros(n_classes, dtype=np.int64)
        for i in rng.integers(n_classes, size=number_of_object):
            counts[i] += 1
    else:
        counts = np.array(fixed_counts, dtype=np.int64)
    events = int(counts.sum()) + 1
    perms = np.zeros((events, n_classes, tries), dtype=np.int32)
    for e in range(events):
        for c in range(n_classes):
            p = rng.permutation(list_lens[c])
            k = min(tries, len(p))
            perms[e, c, :k] = p[:k]
            perms[e, c, k:] = -1
    return Schedule(counts=counts, perms=perms)


# ------------------------------------------------------------------------------------------- whole cases
def load_config(task):
    import os
    import yaml
    name = "KITTI.yaml" if task == "od" else "semantic-kitti.yaml"
    with open(os.path.join(os.path.dirname(__file__), "config", name), "r") as f:
        return yaml.safe_load(f)


@dataclass
class Case:
    """All inputs of one scan's augmentation, in the reference's layouts."""
    task: str
    config: dict
    pcl5: np.ndarray               # N x 5 float64: x y z intensity label  (dataset __getitem__, od/ds:62-66)
    box_lines: list                # scene annotation lines
    db: dict                       # {class: [(name, {'pcl','anno'})]}
    schedule: Schedule
    maps: dict | None = None       # OD {'Road': {...}, 'Sidewalk': {...}}
    map_data: dict | None = None   # semseg {'map','move'}
    pose: np.ndarray | None = None
    cars: list = field(default_factory=list)


_DB_CACHE = {}


def make_case(task, seed, shape: ScanShape = KITTI_SHAPE, n_cars=None, n_per_class=100, number_of_object=None,
              counts=None, classes=None, db_seed=0, tilt=True, tries=100, obj_range=(5.0, 35.0)):
    """Seeded scan + boxes + maps + cut-object DB + schedule for ``task`` in {'od', 'ss'}."""
    semseg = task == "ss"
    cfg = load_config(task)
    if classes is not None:
        cfg["insertion"]["classes"] = list(classes)
    classes = cfg["insertion"]["classes"]
    rng = np.random.default_rng([seed, 5])
    if n_cars is None:
        n_cars = int(rng.integers(0, 16))
    cars = make_scene_cars(seed, n_cars)
    pcl, labels = make_scan(seed, shape, cars)
    pcl5 = np.hstack((pcl, (labels & 0xFFFF).reshape(-1, 1)))         # float64, as np.hstack promotes (od/ds:66)
    key = (task, db_seed, tuple(classes), n_per_class, shape.beams, shape.az_steps, tuple(obj_range))
    if key not in _DB_CACHE:
        _DB_CACHE[key] = make_object_db(1000 + db_seed, classes, n_per_class, semseg, shape, obj_range)
    db = _DB_CACHE[key]
    if number_of_object is None:
        number_of_object = int(np.sum(counts)) if counts is not None else cfg["insertion"]["number_of_object"]
    cfg["insertion"]["number_of_object"] = number_of_object
    sched = make_schedule(seed, len(classes), number_of_object, [len(db[c]) for c in classes], tries=tries,
                          random_counts=counts is None, fixed_counts=counts)
    case = Case(task, cfg, pcl5, make_scene_box_lines(cars, semseg), db, sched, cars=cars)
    if semseg:
        case.pose = make_pose(seed, tilt)
        case.map_data = make_ss_map(case.pose)
    else:
        road, side = make_od_maps()
        case.maps = {"Road": road, "Sidewalk": side}
    return case


def array_digest(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def case_digest(case: Case):
    parts = [case.pcl5, case.schedule.counts, case.schedule.perms]
    for c in case.config["insertion"]["classes"]:
        for name, s in case.db[c]:
            parts.append(s["pcl"])
            parts.append(np.frombuffer(str(s["anno"]).encode(), dtype=np.uint8))
    for l in case.box_lines:
        parts.append(np.frombuffer(l.encode(), dtype=np.uint8))
    if case.maps is not None:
        parts += [case.maps["Road"]["map"], case.maps["Sidewalk"]["map"]]
    if case.map_data is not None:
        parts += [case.map_data["map"], case.map_data["move"], case.pose]
    return array_digest(*parts)
