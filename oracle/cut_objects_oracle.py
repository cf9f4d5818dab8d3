"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's cut-object database builders.

  od/co = object_detection/cut_object/object_cut_out.py      od/cu = object_detection/cut_object/cutout.py
  ss/co = semantic_segmentation/cut_object/cut_out.py        ss/fo = semantic_segmentation/cut_object/filter_objects.py

Pinned by ``tests/golden/cut_objects_{od,ss}.npz`` = what the UNMODIFIED scripts wrote for a seeded synthetic dataset
(``oracle/make_golden.py``).  Only ``tests/`` and CPU baseline legs may import this module.
"""
from __future__ import annotations

import copy
import math

import numpy as np
from scipy.spatial.transform import Rotation as R

from .real3d_oracle import cut_bounding_box


def _dictionary(a):
    """od/co:20-25, ss/co:17-22."""
    return {'center': {'x': a[0][0], 'y': a[0][1], 'z': a[0][2]},
            'rotation': {'x': a[1][0], 'y': a[1][1], 'z': a[1][2], 'w': a[1][3]},
            'length': a[2][0], 'width': a[2][1], 'height': a[2][2]}


def _yaw_quat(z_rot):
    rot_matrix = [[math.cos(z_rot), -1 * math.sin(z_rot), 0], [math.sin(z_rot), math.cos(z_rot), 0], [0, 0, 1]]
    return R.from_matrix(rot_matrix).as_quat()


# ------------------------------------------------------------------------------------------------ camera (od/cu)
def read_calib(calib_file):
    """od/cu:35-52 — float32 matrices from lines 2..5 of a KITTI calib file."""
    with open(calib_file) as f:
        lines = f.readlines()
    row = lambda i: np.array(lines[i].strip().split(' ')[1:], dtype=np.float32)
    return {'P2': row(2).reshape(3, 4), 'R0': row(4).reshape(3, 3), 'Tr_velo2cam': row(5).reshape(3, 4)}


def fov_flag(xyz, calib, img_shape):
    """od/cu:73-112: lidar_to_rect, rect_to_img, get_fov_flag (xyz float64 N x 3)."""
    hom = np.hstack((xyz, np.ones((xyz.shape[0], 1), dtype=np.float32)))
    pts_rect = np.dot(hom, np.dot(calib['Tr_velo2cam'].T, calib['R0'].T))                      # od/cu:78-81
    rect_hom = np.hstack((pts_rect, np.ones((pts_rect.shape[0], 1), dtype=np.float32)))
    pts_2d_hom = np.dot(rect_hom, calib['P2'].T)                                               # od/cu:94-96
    pts_img = (pts_2d_hom[:, 0:2].T / rect_hom[:, 2]).T                                        # od/cu:97
    depth = pts_2d_hom[:, 2] - calib['P2'].T[3, 2]                                             # od/cu:98
    flag = np.logical_and(pts_img[:, 0] >= 0, pts_img[:, 0] < img_shape[1])
    flag &= np.logical_and(pts_img[:, 1] >= 0, pts_img[:, 1] < img_shape[0])
    return np.logical_and(flag, depth >= 0)                                                    # od/cu:103-112


# ------------------------------------------------------------------------------------------------ OD (od/co:78-168)
def od_boxes(annotation):
    """od/co:93-139 for one label line: None if the line is skipped, else (class, base box, expanded box,
    corrected x, corrected y)."""
    items = annotation.split(' ')
    sample_class = items[0]
    sample_occluded = int(items[2])
    h, w, l = float(items[8]), float(items[9]), float(items[10])
    x, y, z = float(items[11]), float(items[12]), float(items[13])
    corrected_x, corrected_y, corrected_z = float(z) + 0.27, float(x) * -1, float(y) * -1 - 0.08
    q = _yaw_quat(float(items[14]) * -1)
    base = _dictionary([[corrected_x, corrected_y, corrected_z], [q[0], q[1], q[2], q[3]], [w + 0.2, l + 0.2, h + 0.1]])
    expand = copy.deepcopy(base)
    expand['length'] += 0.2
    expand['width'] += 0.2
    expand['height'] += 0.2
    return sample_class, sample_occluded, base, expand, corrected_x, corrected_y


def cut_objects_od(points5, label_lines, calib, img_shape, config, frame):
    """od/co:78-168 for one frame.  ``points5``: N x 5 float64 (x, y, z, intensity, label).  Returns the list of
    (file name without '.npz', annotation line, pcl M x 5) the script would save, in order."""
    classes = config['insertion']['classes']
    classes_count = np.zeros(len(classes))
    out = []
    for annotation in label_lines:
        if len(annotation) == 0:
            break
        if not (annotation.split(' ')[0] in classes):
            continue
        sample_class, occluded, base, expand, cx, cy = od_boxes(annotation)
        if occluded != 0:                                                                      # od/co:104-107
            continue
        expand_bbox = cut_bounding_box(points5, expand)
        if len(expand_bbox) != int(fov_flag(expand_bbox[:, 0:3], calib, img_shape).sum()):     # od/co:144
            continue
        classes_count[classes.index(sample_class)] += 1
        bb = cut_bounding_box(points5, base)
        for name in ('Road', 'Parking', 'Sidewalk'):                                           # od/co:150-152
            bb = bb[bb[:, 4] != config['labels'][name]]
        bb = bb[:, 0:4]
        if len(bb) < config['insertion']['min_points'][sample_class]:
            continue
        bb = np.hstack((bb, np.ones((len(bb), 1))))
        name = (f'{config["insertion"]["labels_shortcut"][sample_class]}{frame}_'
                f'{int(classes_count[classes.index(sample_class)])}_{int(np.sqrt(cx ** 2 + cy ** 2))}_m')
        out.append((sample_class, name, annotation, bb))
    return out


# ------------------------------------------------------------------------------------------------ semseg (ss/co:69-157)
def cut_objects_ss(points5, anno_lines, config, sequence, frame):
    """ss/co:96-157 for one frame: list of (label folder, file name without '.npz', annotation line, pcl M x 5)."""
    classes = config['insertion']['classes']
    classes_count = np.zeros(len(classes))
    out = []
    for annotation in anno_lines:
        if len(annotation) == 0:
            break
        items = annotation.split(' ')
        cls = int(items[0])
        if not (cls in classes):
            continue
        classes_count[classes.index(cls)] += 1
        x, y, z = float(items[1]), float(items[2]), float(items[3])
        height, width, length = float(items[4]), float(items[6]), float(items[5])
        q = _yaw_quat(float(items[7]))
        box = _dictionary([[x, y, z], [q[0], q[1], q[2], q[3]], [width, length, height]])
        bb = cut_bounding_box(points5, box)
        bb = bb[bb[:, 4] == cls]                                                               # ss/co:143
        if len(bb) < config['insertion']['min_points'][cls]:
            continue
        shortcut = config['insertion']['labels_shortcut'][cls]
        name = f'{shortcut}{sequence}-{frame}_{int(classes_count[classes.index(cls)]):02d}_{int(np.sqrt(x ** 2 + y ** 2)):03d}_m'
        out.append((config['labels'][cls], name, annotation, bb))
    return out


def filter_objects(samples):
    """ss/fo:88-115 for the samples of ONE class and ONE distance bucket: ``samples`` = list of (name, annotation,
    number of points); returns the names the script deletes (fewer points than the average of their 1-degree yaw bin)."""
    total, number = np.zeros(360), np.zeros(360)
    rot = lambda anno: int(np.rad2deg(float(str(anno).split(' ')[7])) + 180)
    for _, anno, n in samples:
        total[rot(anno)] += n
        number[rot(anno)] += 1
    with np.errstate(divide='ignore', invalid='ignore'):
        avg = np.where(number != 0, total / number, np.inf)
    return [name for name, anno, n in samples if avg[rot(anno)] > n]
