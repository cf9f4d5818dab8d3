"""TEST INFRASTRUCTURE ONLY — numpy CPU restatement of Real3D-Aug's per-scan hot path (the parity oracle).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / the timed CPU baseline.  The product
(``pcl_augmentation_b200``) never imports it and has no CPU fallback.

Pinning: this restatement is checked against the UNMODIFIED reference (imported from /root/reference through
``oracle/shim.py``) by ``tests/test_oracle_vs_reference.py`` (runs where /root/reference exists) and against
the committed golden fixtures ``tests/golden/*.npz`` that ``oracle/make_golden.py`` produced by running the
reference's own functions and its ``insertion.py`` ``__main__`` on seeded synthetic inputs.  The reference has no
tests or golden vectors of its own.  One step — ``skimage.morphology.closing`` — lives in an
un-vendored, un-pinned third-party dependency (scikit-image, absent from this image): its parity is pinned by
definition (binary 5x3 dilate-then-erode, out-of-image ignored), see ``oracle/shim.py``.

Abbreviations for citations (paths under /root/reference):
  od/ins = object_detection/Real3DAug/insertion.py          ss/ins = semantic_segmentation/Real3DAug/insertion.py
  od/fs  = object_detection/Real3DAug/tools/find_spot.py    ss/fs  = semantic_segmentation/Real3DAug/tools/find_spot.py
  cb     = */Real3DAug/tools/cut_bbox.py                    cl     = */Real3DAug/tools/closing.py
  od/ds  = object_detection/Real3DAug/tools/datasets.py     ss/ds  = semantic_segmentation/Real3DAug/tools/datasets.py

``mode='cumulative'`` follows the reference literally (candidate k = k successive in-place +step rotations, carried
z shift).  ``mode='closed'`` is the stateless closed form the CUDA path implements (candidate k = one rotation by
k*step of the original sample; box yaw matrix = R0 @ Rz(k*step)); both give the same feasible sets / choices and xyz
within ~1e-12 m (tests/test_oracle.py).
"""
from __future__ import annotations

import copy
import math

import numpy as np
from scipy.spatial.transform import Rotation as R

NUMROW = 112                 # od/ins:21
NUMCOLUMN = 360 * 4          # od/ins:22
MAX_NUM_TRIES = 100          # od/ins:24
ROAD_INDEXES = [40, 44, 48]  # od/fs:14 (Road, Parking, Sidewalk)


# ======================================================================================= A1-A3 projection
def add_space_for_spherical(point_cloud):
    """od/ins:55-65 — N x 5 -> N x 9 (x, y, z, r, az, el, intensity, label, pix_id), unset = -1."""
    out = np.ones((len(point_cloud), 9)) * -1
    out[:, 0:3] = point_cloud[:, 0:3]
    out[:, 6:8] = point_cloud[:, 3:5]
    return out


def fill_spherical(point_cloud):
    """od/ins:68-82 — in-place r / azimuth / elevation; returns (pc, max_el, min_el)."""
    point_cloud[:, 3] = np.sqrt(point_cloud[:, 0] ** 2 + point_cloud[:, 1] ** 2 + point_cloud[:, 2] ** 2)
    point_cloud[:, 4] = np.arctan2(point_cloud[:, 1], point_cloud[:, 0]) + np.pi
    point_cloud[:, 5] = np.arccos(point_cloud[:, 2] / point_cloud[:, 3])
    return point_cloud, np.max(point_cloud[:, 5]), np.min(point_cloud[:, 5])


def bin_rows_cols(el, az, num_row, num_column, max_el, min_el):
    """od/ins:97-98,105-106 — truncation-toward-zero binning with the -1e-5 elevation offset."""
    d_elevation = (max_el - min_el) / num_row
    d_azimuth = 2 * math.pi / num_column
    rows = np.trunc((el - min_el - 0.00001) / d_elevation).astype(np.int64)
    cols = np.trunc(np.mod(az, 2 * math.pi) / d_azimuth).astype(np.int64)
    return rows, cols


def geometrical_front_view(point_cloud, num_row, num_column, max_elevation_angle, min_elevation_angle,
                           sample=False, numcolumn_global=None):
    """od/ins:85-130 — spherical z-buffer.  ``train`` = min range per pixel (empty = 500), ``label`` = 1 where a
    point landed (empty = -1), pix_id = row * NUMCOLUMN + col written to column 8 (the module global, od/ins:117)."""
    if numcolumn_global is None:
        numcolumn_global = num_column
    label = np.ones((num_row, num_column)) * -1
    train = np.ones((num_row, num_column)) * 500
    if len(point_cloud) == 0:
        return train, label, point_cloud
    rows, cols = bin_rows_cols(point_cloud[:, 5], point_cloud[:, 4], num_row, num_column,
                               max_elevation_angle, min_elevation_angle)
    row_ok = (rows >= 0) & (rows < num_row)
    if sample:
        use = row_ok                                              # od/ins:108-109
    else:
        assert row_ok.all(), "Rows in FoV went something wrong."   # od/ins:111
        use = np.ones(len(rows), dtype=bool)
    assert ((cols[use] >= 0) & (cols[use] < num_column)).all(), "Column in FoV went something wrong."   # od/ins:113
    flat = rows[use] * num_column + cols[use]
    best = np.full(num_row * num_column, np.inf)
    np.minimum.at(best, flat, point_cloud[use, 3])
    hit = np.isfinite(best)
    train.reshape(-1)[hit] = best[hit]
    label.reshape(-1)[hit] = 1
    point_cloud[use, 8] = rows[use] * numcolumn_global + cols[use]
    return train, label, point_cloud


# ============================================================================================ A4 closing
def _shift(a, dr, dc, fill):
    """a shifted so that out[r, c] = a[r + dr, c + dc] (``fill`` outside the image)."""
    h, w = a.shape
    out = np.full_like(a, fill)
    r0, r1 = max(0, -dr), min(h, h - dr)
    c0, c1 = max(0, -dc), min(w, w - dc)
    if r0 < r1 and c0 < c1:
        out[r0:r1, c0:c1] = a[r0 + dr:r1 + dr, c0 + dc:c1 + dc]
    return out


def class_closing(original_label):
    """cl:9-23 — occupancy -> ubyte 0/255 -> closing(rectangle(5, 3)): 5 rows x 3 cols dilate then erode,
    pixels outside the image ignored (== scipy/skimage 'reflect' for min/max filters), no azimuth wrap."""
    occ = np.clip(original_label, 0, 1) > 0
    dil = np.zeros_like(occ)
    for dr in range(-2, 3):
        for dc in range(-1, 2):
            dil |= _shift(occ, dr, dc, False)
    ero = np.ones_like(occ)
    for dr in range(-2, 3):
        for dc in range(-1, 2):
            ero &= _shift(dil, dr, dc, True)
    return (ero * 255).astype(np.uint8)


def smooth_out(original_train, original_label):
    """cl:26-62 — pixels the closing switched on but that hold no point get the mean range of their occupied
    5x3 neighbours (ORIGINAL values, summed in the reference's (drow, dcol) order) and label 1."""
    train = original_train.copy()
    label = original_label.copy()
    closed = class_closing(original_label)
    fill = (closed == 255) & (label != 1)
    occ = original_label == 1
    neighbors = np.zeros(original_label.shape, dtype=np.int64)
    sum_distance = np.zeros(original_label.shape)
    for drow in range(-2, 3):
        for dcolumn in range(-1, 2):
            o = _shift(occ, drow, dcolumn, False)
            t = _shift(original_train, drow, dcolumn, 0.0)
            neighbors += o
            sum_distance = sum_distance + np.where(o, t, 0.0)
    ok = fill & (neighbors > 0)
    train[ok] = sum_distance[ok] / neighbors[ok]
    label[fill] = 1
    return train, label


# ================================================================================================ boxes
def make_dictionary(a, ss=False):
    """od/fs:44-55, ss/fs:15-27 (semseg keeps the class as a 1-element list)."""
    return {'center': {'x': a[0][0], 'y': a[0][1], 'z': a[0][2]},
            'rotation': {'x': a[1][0], 'y': a[1][1], 'z': a[1][2], 'w': a[1][3]},
            'length': a[2][0], 'width': a[2][1], 'height': a[2][2], 'class': a[3] if ss else a[3][0]}


def dictionary2array(d, ss=False):
    """od/fs:58-68, ss/fs:30-39."""
    return [[d['center']['x'], d['center']['y'], d['center']['z']],
            [d['rotation']['x'], d['rotation']['y'], d['rotation']['z'], d['rotation']['w']],
            [d['length'], d['width'], d['height']], d['class'] if ss else [d['class']]]


def _yaw_quat(z_rot):
    rot_matrix = [[math.cos(z_rot), -1 * math.sin(z_rot), 0], [math.sin(z_rot), math.cos(z_rot), 0], [0, 0, 1]]
    return R.from_matrix(rot_matrix).as_quat()


def read_label_line_od(line):
    """od/fs:175-224 — KITTI label_2 line (camera frame) -> lidar-frame box dict (center.z = box bottom)."""
    it = line.split(' ')
    h, w, l = float(it[8]), float(it[9]), float(it[10])
    x, y, z = float(it[11]), float(it[12]), float(it[13])
    q = _yaw_quat(float(it[14]) * -1)
    return make_dictionary([[float(z) + 0.27, float(x) * -1, float(y) * -1 - 0.08], [q[0], q[1], q[2], q[3]],
                            [w + 0.1, l + 0.1, h + 0.1], [it[0]]])


def read_label_line_ss(line):
    """ss/fs:155-189 — ``cls x y z h w l yaw``; dict length = field 6, width = field 5."""
    it = line.split(' ')
    q = _yaw_quat(float(it[7]))
    return make_dictionary([[float(it[1]), float(it[2]), float(it[3])], [q[0], q[1], q[2], q[3]],
                            [float(it[6]), float(it[5]), float(it[4])], [it[0]]], ss=True)


def box_matrix(annotation):
    if '_matrix' in annotation:                      # closed-form candidates carry their matrix (no quat round trip)
        return annotation['_matrix']
    return R.from_quat([annotation['rotation']['x'], annotation['rotation']['y'], annotation['rotation']['z'],
                        annotation['rotation']['w']]).as_matrix()


def cut_bounding_box_mask(point_cloud, annotation, annotation_move=(0, 0, 0), rot_matrix=None):
    """cb:7-68 — strict point-in-OBB test, z measured from the box bottom; same fp expression order."""
    xc = annotation['center']['x'] - annotation_move[0]
    yc = annotation['center']['y'] - annotation_move[1]
    zc = annotation['center']['z'] - annotation_move[2]
    length, width, height = annotation['length'], annotation['width'], annotation['height']
    m = box_matrix(annotation) if rot_matrix is None else rot_matrix
    px, py, pz = point_cloud[:, 0], point_cloud[:, 1], point_cloud[:, 2]
    a0 = m[0][0] * px + m[1][0] * py + m[2][0] * pz
    a1 = m[0][1] * px + m[1][1] * py + m[2][1] * pz
    a2 = m[0][2] * px + m[1][2] * py + m[2][2] * pz
    mask = a0 < m[0][0] * (xc + m[0][0] * length / 2) + m[1][0] * (yc + m[1][0] * length / 2) + m[2][0] * (zc + m[2][0] * length / 2)
    mask &= a0 > m[0][0] * (xc - m[0][0] * length / 2) + m[1][0] * (yc - m[1][0] * length / 2) + m[2][0] * (zc - m[2][0] * length / 2)
    mask &= a1 < m[0][1] * (xc + m[0][1] * width / 2) + m[1][1] * (yc + m[1][1] * width / 2) + m[2][1] * (zc + m[2][1] * width / 2)
    mask &= a1 > m[0][1] * (xc - m[0][1] * width / 2) + m[1][1] * (yc - m[1][1] * width / 2) + m[2][1] * (zc - m[2][1] * width / 2)
    mask &= a2 < m[0][2] * (xc + m[0][2] * height) + m[1][2] * (yc + m[1][2] * height) + m[2][2] * (zc + m[2][2] * height)
    mask &= a2 > m[0][2] * (xc - m[0][2] * 0) + m[1][2] * (yc - m[1][2] * 0) + m[2][2] * (zc - m[2][2] * 0)
    return mask


def cut_bounding_box(point_cloud, annotation, annotation_move=(0, 0, 0)):
    return point_cloud[cut_bounding_box_mask(point_cloud, annotation, annotation_move)]


# ============================================================================================ placement
def yaw_tables(yaw_steps):
    """cos/sin of k * (360 / yaw_steps) degrees for k = 0..yaw_steps (closed-form candidates)."""
    ang = np.deg2rad(np.arange(yaw_steps + 1) * (360.0 / yaw_steps))
    return np.cos(ang), np.sin(ang)


def rotate_bounding_box(bbox_pcl, annotation, rotation=1, ss=False):
    """od/fs:71-106, ss/fs:42-76 — rotate points and box centre about the SENSOR z-axis, box quat <- R_box . Rz."""
    a = dictionary2array(annotation, ss)
    rotation = np.deg2rad(rotation)
    rot_matrix = R.from_quat(a[1]).as_matrix()
    z_rot_matrix = np.array([[np.cos(rotation), -np.sin(rotation), 0], [np.sin(rotation), np.cos(rotation), 0], [0, 0, 1]])
    a[1] = R.from_matrix(np.dot(rot_matrix, z_rot_matrix)).as_quat()
    position = np.dot(z_rot_matrix, np.array([[a[0][0]], [a[0][1]], [a[0][2]]]))
    a[0][0], a[0][1], a[0][2] = position[0][0], position[1][0], position[2][0]
    bbox_pcl[:, :3] = (z_rot_matrix @ bbox_pcl[:, :3].T).T
    return bbox_pcl, make_dictionary(a, ss)


RADII = []
_r = 0.1
for _ in range(50):
    RADII.append(_r)
    _r += 0.1
RADII_OK = [RADII[i] + 0.1 <= 5 for i in range(50)]      # od/fs:156-160: a hit on the pass that pushes radius > 5 fails
del _r


def road_level(ground_xyz, cx, cy):
    """od/fs:149-164, ss/fs:118-144 — growing-radius search (0.1, 0.2, ... by repeated += 0.1, <= 50 passes) for
    surface points around (cx, cy); returns (mean z, ok).  ``ground_xyz`` = rows of the ORIGINAL scene already
    filtered by surface label and z > -3 (row order preserved, so the mean sums in the reference's order)."""
    if len(ground_xyz) == 0:
        return 0.0, False
    d2 = (ground_xyz[:, 0] - cx) ** 2 + (ground_xyz[:, 1] - cy) ** 2
    for j, radius in enumerate(RADII):
        sel = d2 <= radius ** 2
        if sel.any():
            if not RADII_OK[j]:
                return 0.0, False
            surface = ground_xyz[sel]
            return np.mean(surface, axis=0)[2], True
    return 0.0, False


def _ground_rows_od(original_pcl, road_label):
    g = original_pcl[original_pcl[:, 4] == road_label]
    return g[g[:, 2] > -3][:, :3]


def _ground_rows_ss(original_pcl, ok_surface):
    parts = [original_pcl[original_pcl[:, 4] == s] for s in ok_surface]      # ss/fs:125-131: grouped by label
    g = np.concatenate(parts, axis=0) if parts else original_pcl[:0]
    return g[g[:, 2] > -3][:, :3]


def _annulus(points_xy, cx, cy, reach):
    rho = math.hypot(cx, cy)
    pr2 = points_xy[:, 0] ** 2 + points_xy[:, 1] ** 2
    lo = max(rho - reach, 0.0)
    return (pr2 >= lo * lo) & (pr2 <= (rho + reach) ** 2)


def _box_reach(anno):
    return 0.5 * math.hypot(anno['length'], anno['width']) + 0.25


def on_map_od(sample_pcl, map_arr, map_move):
    """od/fs:267-279 — every in-map object point must sit on a cell == 1; no in-map point -> reject."""
    gx = sample_pcl[:, 0] - map_move[0]
    gy = sample_pcl[:, 1] - map_move[1]
    inmap = ~((gx < 0) | (gx >= map_arr.shape[0]) | (gy < 0) | (gy >= map_arr.shape[1]))
    if not inmap.any():
        return False
    cells = map_arr[gx[inmap].astype(np.int64), gy[inmap].astype(np.int64)]
    return bool((cells == 1).all())


def on_map_ss(sample_pcl, map_arr, map_move, transformation_matrix, ok_map_surface):
    """ss/fs:235-248 — world = T . [x y z 1] - move, astype(int) (trunc toward zero); every in-map cell value must
    be in ``placement[class]``; no in-map point -> accept."""
    gp = np.hstack((sample_pcl[:, :3], np.ones((len(sample_pcl), 1)))).T
    gp = transformation_matrix @ gp
    gp = gp - map_move
    gp = gp.astype(int)
    gp = gp[:, gp[0, :] < len(map_arr)]
    gp = gp[:, gp[0, :] > -1]
    gp = gp[:, gp[1, :] < len(map_arr[0])]
    gp = gp[:, gp[1, :] > -1]
    cells = map_arr[gp[0], gp[1]]
    return bool(np.isin(cells, ok_map_surface).all())


def collide_od(scene_pcl, scene_boxes, sample_pcl, sample_anno, rot_matrix=None):
    """od/fs:109-135 — obstacle scene points (working label == 1; Pedestrian: only z >= box bottom + 0.1) inside the
    candidate box, or any object point inside an existing box."""
    m = cut_bounding_box_mask(scene_pcl, sample_anno, rot_matrix=rot_matrix)
    m &= scene_pcl[:, 7] == 1
    if sample_anno['class'] == 'Pedestrian':
        m &= scene_pcl[:, 2] >= sample_anno['center']['z'] + 0.1
    if m.any():
        return True
    for b in scene_boxes:
        if cut_bounding_box_mask(sample_pcl, b).any():
            return True
    return False


def collide_ss(scene_pcl, scene_boxes, sample_pcl, sample_anno, ok_surface, rot_matrix=None):
    """ss/fs:79-104 — any scene point inside the candidate box whose label is not an allowed surface label."""
    m = cut_bounding_box_mask(scene_pcl, sample_anno, rot_matrix=rot_matrix)
    m &= ~np.isin(scene_pcl[:, 7], ok_surface)
    if m.any():
        return True
    for b in scene_boxes:
        if cut_bounding_box_mask(sample_pcl, b).any():
            return True
    return False


def _find_possible_places(point_cloud, scene_annotation, sample_pcl, sample_annotation, ground, on_map, collide,
                          ss, yaw_steps, mode, fast):
    """Shared 360-yaw loop of od/fs:263-300 and ss/fs:231-269."""
    out_pcl, out_anno, out_rot = [], [], []
    step = 360.0 / yaw_steps
    if fast:
        reach = _box_reach(sample_annotation)
        cx0, cy0 = sample_annotation['center']['x'], sample_annotation['center']['y']
        point_cloud = point_cloud[_annulus(point_cloud[:, :2], cx0, cy0, reach)]
        ground = ground[_annulus(ground[:, :2], cx0, cy0, 5.2)]
    if mode == 'closed':
        cos_t, sin_t = yaw_tables(yaw_steps)
        pcl0 = sample_pcl.copy()
        anno0 = copy.deepcopy(sample_annotation)
        m0 = box_matrix(anno0)
        c0x, c0y, z_box0 = anno0['center']['x'], anno0['center']['y'], anno0['center']['z']
        dz = 0.0                                        # carried z shift (od/fs:167-168 is in place)
    for rot in range(1, yaw_steps + 1):
        if mode == 'cumulative':
            sample_pcl, sample_annotation = rotate_bounding_box(sample_pcl, sample_annotation, step, ss)
            rot_matrix = None
        else:
            c, s = cos_t[rot], sin_t[rot]
            sample_pcl = pcl0.copy()
            sample_pcl[:, 0] = c * pcl0[:, 0] - s * pcl0[:, 1]
            sample_pcl[:, 1] = s * pcl0[:, 0] + c * pcl0[:, 1]
            sample_pcl[:, 2] = pcl0[:, 2] + dz
            rot_matrix = np.array([[m0[0][0] * c - m0[1][0] * s, -(m0[0][0] * s + m0[1][0] * c), 0.0],
                                   [m0[1][0] * c + m0[0][0] * s, m0[0][0] * c - m0[1][0] * s, 0.0],
                                   [0.0, 0.0, 1.0]])
            sample_annotation = copy.deepcopy(anno0)
            sample_annotation['center']['x'] = c * c0x - s * c0y
            sample_annotation['center']['y'] = s * c0x + c * c0y
            sample_annotation['center']['z'] = z_box0 + dz
            q = R.from_matrix(rot_matrix).as_quat()
            sample_annotation['rotation'] = {'x': q[0], 'y': q[1], 'z': q[2], 'w': q[3]}
            sample_annotation['_matrix'] = rot_matrix
        if not on_map(sample_pcl):
            continue
        level, near_road = road_level(ground, sample_annotation['center']['x'], sample_annotation['center']['y'])
        if not near_road:
            continue
        if mode == 'cumulative':
            z_move = level - sample_annotation['center']['z']             # od/fs:164-168
            sample_pcl[:, 2] += z_move
            sample_annotation['center']['z'] = level
        else:
            dz = level - z_box0
            sample_pcl[:, 2] = pcl0[:, 2] + dz
            sample_annotation['center']['z'] = level
        if not collide(point_cloud, scene_annotation, sample_pcl, sample_annotation, rot_matrix):
            out_pcl.append(copy.deepcopy(sample_pcl))
            out_anno.append(copy.deepcopy(sample_annotation))
            out_rot.append(rot)
    return out_pcl, out_anno, out_rot


def find_possible_places_od(point_cloud, scene_annotation, sample_data, map_data, original_pcl, config,
                            yaw_steps=360, mode='cumulative', fast=True):
    """od/fs:227-304."""
    sample_pcl = sample_data['pcl']
    sample_pcl[:, 4] = 1
    sample_annotation = read_label_line_od(sample_data['anno'].item())
    map_arr = map_data['map']
    map_move = np.array([map_data['min_x'], map_data['min_y']])
    ground = _ground_rows_od(original_pcl, config['labels']['Road'])
    return _find_possible_places(
        point_cloud, scene_annotation, sample_pcl, sample_annotation, ground,
        lambda p: on_map_od(p, map_arr, map_move),
        lambda sc, boxes, sp, sa, rm: collide_od(sc, boxes, sp, sa, rm),
        False, yaw_steps, mode, fast)


def find_possible_places_ss(point_cloud, scene_annotation, sample_data, map_arr, map_move, original_pcl,
                            transformation_matrix, config, yaw_steps=360, mode='cumulative', fast=True):
    """ss/fs:192-273."""
    sample_pcl = sample_data['pcl']
    sample_annotation = read_label_line_ss(sample_data['anno'].item())
    ok_map_surface = config['insertion']['placement'][int(sample_annotation['class'][0])]
    ok_surface = []
    for map_surface in ok_map_surface:
        ok_surface = ok_surface + config['insertion']['placement_labels'][map_surface]
    ground = _ground_rows_ss(original_pcl, ok_surface)
    return _find_possible_places(
        point_cloud, scene_annotation, sample_pcl, sample_annotation, ground,
        lambda p: on_map_ss(p, map_arr, map_move, transformation_matrix, ok_map_surface),
        lambda sc, boxes, sp, sa, rm: collide_ss(sc, boxes, sp, sa, ok_surface, rm),
        True, yaw_steps, mode, fast)


def addjust_map_2(map_data, point_cloud, transformation_matrix, road_indexes=ROAD_INDEXES):
    """ss/ins:202-224 — cells (value != 0) holding scene points with z < 1.5 and a non-ground label become 4."""
    map_arr = map_data['map']
    map_move = map_data['move']
    pc = point_cloud[point_cloud[:, 2] < 1.5]
    for i in road_indexes:
        pc = pc[pc[:, 7] != i]
    hom = np.hstack((pc[:, :3], np.ones((len(pc), 1)))).T
    hom = transformation_matrix @ hom
    hom = (hom - map_move).astype(int)
    ix, iy = hom[0], hom[1]
    sel = map_arr[ix, iy] != 0
    map_arr[ix[sel], iy[sel]] = 4
    return map_arr, map_move


# ============================================================================================ occlusion
def occlude(scene_pcl, scene_train, sample_pcl5, num_row, num_column, max_elevation, min_elevation,
            numcolumn_global=None):
    """od/ins:472-501 — project the candidate with the SCENE's elevation range, smooth it, compare the smoothed
    images with strict <, drop every scene point of a visible pixel, keep the object points of visible pixels in
    (pix_id, original index) order.  Returns (scene_keep_mask, visible_sample V x 9, vis_px bool H x W)."""
    if numcolumn_global is None:
        numcolumn_global = num_column
    sample_pcl = add_space_for_spherical(sample_pcl5)
    sample_pcl, _, _ = fill_spherical(sample_pcl)
    sample_train, sample_label, sample_pcl = geometrical_front_view(
        sample_pcl, num_row, num_column, max_elevation, min_elevation, sample=True, numcolumn_global=numcolumn_global)
    sample_train, sample_label = smooth_out(sample_train, sample_label)
    vis = sample_train < scene_train
    rr, cc = np.where(vis)
    vis_ids = rr * numcolumn_global + cc
    scene_keep = ~np.isin(scene_pcl[:, 8], vis_ids)
    spix = sample_pcl[:, 8]
    sel = np.isin(spix, vis_ids)                      # pix_id -1 never matches
    idx = np.nonzero(sel)[0]
    order = np.argsort(spix[idx], kind='stable')
    return scene_keep, sample_pcl[idx[order]], vis


# ======================================================================================= A12-A14 driver
def create_annotation_line(original_string, new_annotation_dict, rotation):
    """od/ins:227-265 — KITTI line of an inserted object (lidar->camera offsets -0.27 / -0.08)."""
    rotation = np.deg2rad(rotation)
    cx, cy, cz = (new_annotation_dict['center'][k] for k in 'xyz')
    items = original_string.item().split(' ')
    ry = float(items[14]) - rotation
    if ry < -np.pi:
        ry += 2 * np.pi
    elif ry > np.pi:
        ry -= 2 * np.pi
    assert -np.pi <= ry <= np.pi
    alpha = (np.arctan2((cy * -1), cx - 0.27) * -1) + ry
    if alpha < -np.pi:
        alpha += 2 * np.pi
    elif alpha > np.pi:
        alpha -= 2 * np.pi
    assert -np.pi <= alpha <= np.pi
    return (f"{new_annotation_dict['class']} {items[1]} 3 {alpha:.02f} {items[4]} {items[5]} {items[6]} {items[7]} "
            f"{items[8]} {items[9]} {items[10]} {(cy * -1):.02f} {(cz * -1) - 0.08:.02f} {cx - 0.27:.02f} {ry:.02f}\n")


def remove_space_for_spherical(point_cloud):
    """od/ds:97-109, ss/ds:93-106 — N x 9 -> (N x 4 xyz+intensity, N x 1 label)."""
    n = len(point_cloud)
    pcl = np.ones((n, 4)) * -1
    labels = np.zeros((n, 1))
    if n:
        pcl[:, 0:3] = point_cloud[:, 0:3]
        pcl[:, 3] = point_cloud[:, 6]
        labels[:, 0] = point_cloud[:, 7]
    return pcl, labels


class FreshDict(dict):
    """NpzFile stand-in: each ``[...]`` returns a fresh copy (an NpzFile re-reads the array per access)."""

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        return v.copy() if isinstance(v, np.ndarray) else v


def augment_scan(task, scene_pcl5, scene_box_lines, db, counts, perms, config, *, maps=None, map_data=None,
                 transform_matrix=None, num_row=NUMROW, num_column=NUMCOLUMN, yaw_steps=360, mode='cumulative',
                 fast=True, max_tries=MAX_NUM_TRIES, trace=None):
    """The per-scan ``while`` loop of od/ins:351-628 (task='od') / ss/ins:355-599 (task='ss').

    scene_pcl5  N x 5 float64 (x, y, z, intensity, label) as the dataset adapter returns it (od/ds:62-66)
    db          {class: [(name, {'pcl', 'anno'}), ...]} sorted by name (the sorted glob of sample_path/<class>)
    counts      per-class objects to insert (``generate_seed`` od/ins:171-187)
    perms       int [events, classes, tries] pre-drawn ``random.shuffle`` results (first ``tries`` indices)
    maps        OD: {'Road': map_data_road, 'Sidewalk': map_data_sidewalk}; map_data: semseg {'map', 'move'}
    Returns dict(scene N' x 9, added V x 9, inserted [(name, rot, class)], lines [str], keep_orig bool N,
    n_events).
    """
    ss = task == 'ss'
    classes = config['insertion']['classes']
    remaining = np.array(counts, dtype=np.float64).copy()
    inserted_class = None
    for i in range(len(remaining)):
        if remaining[i] > 0:
            inserted_class = classes[i]
            break
    scene_pcl = np.array(scene_pcl5, dtype=np.float64, copy=True)
    if not ss:
        scene_pcl[scene_pcl[:, 4] != config['labels']['Road'], 4] = 1          # od/ins:353-355
    original_pcl = copy.deepcopy(scene_pcl)
    read_line = read_label_line_ss if ss else read_label_line_od
    scene_annotation = [read_line(l) for l in scene_box_lines]
    n0 = len(scene_pcl)
    scene_pcl = add_space_for_spherical(scene_pcl)
    orig_index = np.arange(n0, dtype=np.int64)            # bookkeeping only: which original rows survive
    lines, inserted = [], []
    all_visible_parts = np.zeros((0, 9))
    SAMPLE_TIMEOUT = False
    unplaceble_samples = []
    event = 0
    start_index = end_index = 0

    while np.max(remaining) > 0:
        scene_pcl, max_elevation, min_elevation = fill_spherical(scene_pcl)
        scene_train, scene_label, scene_pcl = geometrical_front_view(scene_pcl, num_row, num_column, max_elevation,
                                                                     min_elevation)
        scene_train, scene_label = smooth_out(scene_train, scene_label)
        if ss:
            map_arr, map_move = addjust_map_2(FreshDict(map_data), scene_pcl, transform_matrix,
                                              config['insertion'].get('road_indexes', ROAD_INDEXES))   # the name ss/ins:209 reads
        scene_pcl_backup = scene_pcl.copy()
        orig_index_backup = orig_index.copy()
        if trace is not None:
            trace.append(('slot', float(max_elevation), float(min_elevation), len(scene_pcl)))
        for i in range(len(remaining)):
            if remaining[i] > 0:
                if inserted_class != classes[i]:
                    SAMPLE_TIMEOUT = False
                inserted_class = classes[i]
                break
        ci = classes.index(inserted_class)
        base_list = db[inserted_class]
        sample_list = list(range(len(base_list)))
        match_find = False
        while not match_find:
            if not SAMPLE_TIMEOUT:
                head = [int(v) for v in perms[event][ci] if v >= 0]
                head_set = set(head)
                rest = [j for j in range(len(base_list)) if j not in head_set]
                sample_list = head + rest
                event += 1
                start_index, end_index = 0, max_tries
            else:
                start_index += max_tries
                end_index += max_tries
                if end_index > len(sample_list):
                    end_index = len(sample_list)
            for s_index in range(start_index, end_index):
                object_name, sample = base_list[sample_list[s_index]]          # IndexError as in od/ins:410
                if match_find:
                    break
                if object_name in unplaceble_samples:
                    if s_index == end_index - 1:
                        remaining[ci] -= 1
                        match_find = True
                        break
                    continue
                unplaceble = True
                sample_data = FreshDict(sample)
                if ss:
                    poss = find_possible_places_ss(scene_pcl, scene_annotation, sample_data, map_arr, map_move,
                                                   original_pcl, transform_matrix, config, yaw_steps, mode, fast)
                else:
                    placement = config['insertion']['placement'][inserted_class]
                    assert placement in ('Road', 'Sidewalk'), f'unrecognized placement area for {inserted_class}'
                    poss = find_possible_places_od(scene_pcl, scene_annotation, sample_data, maps[placement],
                                                   original_pcl, config, yaw_steps, mode, fast)
                possible_sample_pcl, possible_sample_annotation, possible_rotation = poss
                if trace is not None:
                    trace.append(('try', object_name, list(possible_rotation)))
                if len(possible_sample_pcl) == 0 and unplaceble:
                    unplaceble_samples.append(object_name)
                for sample_index in range(len(possible_sample_pcl)):
                    sample_annotation = possible_sample_annotation[sample_index]
                    sample_rotation = possible_rotation[sample_index]
                    keep, visible_sample, _ = occlude(scene_pcl_backup, scene_train, possible_sample_pcl[sample_index],
                                                      num_row, num_column, max_elevation, min_elevation)
                    scene_pcl = scene_pcl_backup[keep]                      # od/ins:472,491 (persists on failure!)
                    orig_index = orig_index_backup[keep]
                    if trace is not None:
                        trace.append(('cand', object_name, sample_rotation, len(visible_sample)))
                    min_pts = config['insertion']['min_points'][inserted_class]
                    if len(visible_sample) == 0 or len(visible_sample) < min_pts:
                        pass
                    else:
                        inserted.append((object_name, sample_rotation, inserted_class))
                        match_find = True
                        SAMPLE_TIMEOUT = False
                        unplaceble = False
                        scene_pcl = np.append(scene_pcl, visible_sample, axis=0)
                        orig_index = np.append(orig_index, np.full(len(visible_sample), -1, dtype=np.int64))
                        remaining[ci] -= 1
                        all_visible_parts = np.append(all_visible_parts, visible_sample, axis=0)
                        if not ss:
                            # the reference's step is 1 degree, so its rotation index IS the angle (od/ins:553)
                            lines.append(create_annotation_line(sample_data['anno'], sample_annotation,
                                                                sample_rotation * (360.0 / yaw_steps)))
                        if np.max(remaining) > 0:
                            scene_annotation = scene_annotation + [sample_annotation]
                        break
                    if sample_index == len(possible_sample_pcl) - 1 and unplaceble:
                        unplaceble_samples.append(object_name)
                if (s_index == len(sample_list) - 1 or s_index == 3 * max_tries) and not match_find:
                    remaining[ci] = 0
                    SAMPLE_TIMEOUT = True
                if s_index == end_index - 1 and not match_find:
                    remaining[ci] -= 1
                    if remaining[ci] <= 0:
                        match_find = True
                    break
    keep_orig = np.zeros(n0, dtype=bool)
    keep_orig[orig_index[orig_index >= 0]] = True
    return {'scene': scene_pcl, 'added': all_visible_parts, 'inserted': inserted, 'lines': lines,
            'keep_orig': keep_orig, 'n_events': event, 'boxes': scene_annotation}


def save_arrays(task, result):
    """od/ds:76-95, ss/ds:72-91 — what ``save_data`` writes: velodyne f32 N' x 4, (semseg) labels u32, check f32."""
    pcl, labels = remove_space_for_spherical(result['scene'])
    added, added_labels = remove_space_for_spherical(result['added'])
    out = {'velodyne': pcl.astype(np.float32)}
    if task == 'ss':
        out['labels'] = labels.astype(np.uint32)
        out['check'] = np.hstack((added, added_labels)).astype(np.float32)
    else:
        out['check'] = added.astype(np.float32)
    return out
