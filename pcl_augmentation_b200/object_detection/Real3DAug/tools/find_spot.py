"""Drop-in for the reference's object-detection ``tools/find_spot.py``: same function names, argument meaning, in-place
mutation and return values; the point work runs in CUDA (``libreal3d_b200.so``), the few scalars per box on the host.

  make_dictionary :44, dictionary2array :58, rotate_bounding_box :71, check_bounding_box :109, correct_height :138,
  read_label_line :175, find_possible_places :227   (line numbers of the reference file)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.spatial.transform import Rotation as R

from .... import _lib
from .... import boxes as _bx
from ....engine import Real3DEngine, ScanInput
from ....ops import _dev, _stream, cut_bounding_box_mask
from .cut_bbox import cut_bounding_box

ROAD_INDEXES = [40, 44, 48]     # Road, Parking, Sidewalk

RED = '\033[91m'
YELLOW = '\033[93m'
GREEN = '\033[92m'
DEFAULT = '\033[0m'


def make_dictionary(annotation_array):
    return _bx.make_dictionary(annotation_array, ss=False)


def dictionary2array(annotation_dictionary):
    return _bx.dictionary2array(annotation_dictionary, ss=False)


def read_label_line(line):
    return _bx.read_label_line_od(line)


def _transform_in_place(rows, cos_t, sin_t, dz):
    """x' = c x - s y, y' = s x + c y, z' = z + dz on the first three columns of ``rows`` (CUDA), written back in place."""
    if len(rows) == 0:
        return rows
    lib = _lib.load()
    d = _dev(rows, np.float64)
    _lib.check(lib.r3d_transform_points(d.data_ptr(), len(rows), rows.shape[1], float(cos_t), float(sin_t), float(dz),
                                        _stream()), "transform_points")
    rows[:, :3] = d[:, :3].cpu().numpy()
    return rows


def _rotate_annotation(annotation, rotation, ss):
    """Box centre and quaternion of rotate_bounding_box (od/fs:80-96): a handful of scalars, done on the host."""
    a = _bx.dictionary2array(annotation, ss)
    rotation = np.deg2rad(rotation)
    rot_matrix = R.from_quat(a[1]).as_matrix()
    z_rot_matrix = np.array([[np.cos(rotation), -np.sin(rotation), 0], [np.sin(rotation), np.cos(rotation), 0], [0, 0, 1]])
    a[1] = R.from_matrix(np.dot(rot_matrix, z_rot_matrix)).as_quat()
    position = np.dot(z_rot_matrix, np.array([[a[0][0]], [a[0][1]], [a[0][2]]]))
    a[0][0], a[0][1], a[0][2] = position[0][0], position[1][0], position[2][0]
    return _bx.make_dictionary(a, ss), np.cos(rotation), np.sin(rotation)


def rotate_bounding_box(bbox_pcl, annotation, rotation=1):
    """Rotate the object's points (in place) and its annotation about the SENSOR z-axis by ``rotation`` degrees."""
    annotation, c, s = _rotate_annotation(annotation, rotation, ss=False)
    _transform_in_place(bbox_pcl, c, s, 0.0)
    return bbox_pcl, annotation


def check_bounding_box(scene_pcl, scene_anno, sample_pcl, sample_anno):
    """True iff the candidate collides with nothing (od/fs:109-135): one device call (r3d_obb_collide) tests the obstacle
    scene points against the candidate box and the object points against every scene box."""
    from ....ops import obb_collide
    return not bool(obb_collide(scene_pcl, scene_anno, sample_pcl, [(1.0, 0.0, 0.0, sample_anno)], mode='od',
                                pedestrian=sample_anno['class'] == 'Pedestrian')[0])


def _road_level(scene_pcl, labels, cx, cy):
    import torch
    lib = _lib.load()
    r2, ok = _bx.search_radii()
    d = _dev(scene_pcl, np.float64)
    scratch = torch.zeros(3, dtype=torch.int64, device="cuda")
    lab = np.asarray(labels, dtype=np.int32)
    level, found = C.c_double(), C.c_int32()
    _lib.check(lib.r3d_road_level(d.data_ptr(), len(scene_pcl), scene_pcl.shape[1], 4, lab.ctypes.data, len(lab),
                                  float(cx), float(cy), r2.ctypes.data, ok.ctypes.data, scratch.data_ptr(),
                                  C.byref(level), C.byref(found), _stream()), "road_level")
    return level.value, bool(found.value)


def _correct_height(scene_pcl, sample_pcl, sample_anno, labels, ss):
    a = _bx.dictionary2array(sample_anno, ss)
    road_level, ok = _road_level(scene_pcl, labels, a[0][0], a[0][1])
    if ok:
        z_move = road_level - a[0][2]
        _transform_in_place(sample_pcl, 1.0, 0.0, z_move)            # sample_pcl[:, 2] += z_move (od/fs:167)
        a[0][2] = road_level
    return sample_pcl, _bx.make_dictionary(a, ss), ok


def correct_height(scene_pcl, sample_pcl, sample_anno, config):
    """Put the object on the road level found around its box centre in the ORIGINAL scan (od/fs:138-172)."""
    return _correct_height(scene_pcl, sample_pcl, sample_anno, [config['labels']['Road']], ss=False)


def _float32_exact(a):
    return np.array_equal(a.astype(np.float32).astype(np.float64), a)


def _probe(task, point_cloud, scene_annotation, sample_pcl, anno_str, cls_key, original_pcl, config, maps=None,
           map_data=None, pose=None):
    """Shared body of both find_possible_places: a one-scan, one-object engine probes all yaw candidates."""
    if not _float32_exact(original_pcl[:, :4]):
        raise ValueError("original_pcl must hold float32-exact x, y, z, intensity (as read from velodyne/*.bin)")
    db = {cls_key: [("sample", {"pcl": sample_pcl, "anno": anno_str})]}
    eng = Real3DEngine(task, config, db, max_scans=1, max_points=len(original_pcl),
                       max_inserted=max(4096, len(point_cloud)), max_boxes=max(4, len(scene_annotation) + 1),
                       max_events=2, map_data=map_data)
    try:
        n_cls = len(config['insertion']['classes'])
        scan = ScanInput(xyzi=np.ascontiguousarray(original_pcl[:, :4].astype(np.float32)),
                         labels=original_pcl[:, 4].astype(np.uint32), box_lines=[], box_dicts=list(scene_annotation),
                         counts=np.zeros(n_cls, dtype=np.int32), perms=np.zeros((1, n_cls, eng.max_tries), dtype=np.int32),
                         maps=maps, pose=pose)
        eng.load(eng.stage([scan]))
        flags, boxes, xyz, rots = eng.probe_places(0, 0, point_cloud)
        obj_anno = eng.obj_annos[0]
    finally:
        eng.close()
    return flags, boxes, xyz, rots, obj_anno


def find_possible_places(point_cloud, scene_annotation, sample_data, map_data, original_pcl, config):
    """All feasible yaw placements of a cut object (od/fs:227-304): returns (list of M x 5 point clouds, list of box
    dictionaries, list of rotations 1..360) in rotation order."""
    sample_pcl = sample_data['pcl']
    sample_pcl[:, 4] = 1
    anno_str = sample_data['anno']
    cls = read_label_line(anno_str.item())['class']
    flags, boxes, xyz, rots, obj_anno = _probe('od', point_cloud, scene_annotation, sample_pcl, anno_str, cls,
                                                original_pcl, config, maps={'Road': map_data, 'Sidewalk': map_data})
    output_pcl, output_annotation, output_rotation = [], [], []
    for i, k in enumerate(rots):
        pcl = np.array(sample_pcl, copy=True)
        pcl[:, :3] = xyz[i]
        rec = (*boxes[k], obj_anno['length'], obj_anno['width'], obj_anno['height'])
        output_pcl.append(pcl)
        output_annotation.append(_bx.placed_box_dictionary(rec, cls, ss=False))
        output_rotation.append(k)
    not_on_road = int(np.sum((flags[1:] & 3) != 3))
    object_collision = int(np.sum(((flags[1:] & 3) == 3) & ((flags[1:] & 4) != 0)))
    print(f'From 360 possibilities, {YELLOW}{not_on_road}{DEFAULT} was not on road, {RED}{object_collision}{DEFAULT} has '
          f'collision with another object, and {GREEN}{len(output_pcl)}{DEFAULT} was possible.')
    return output_pcl, output_annotation, output_rotation
