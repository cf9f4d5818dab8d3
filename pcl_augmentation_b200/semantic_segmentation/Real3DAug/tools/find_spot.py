"""Drop-in for the reference's semantic-segmentation ``tools/find_spot.py``: same names / arguments / returns, CUDA inside.

  make_dictionary :15, dictionary2array :30, rotate_bounding_box_2 :42, check_bounding_box :79, correct_height :107,
  read_label_line :155, find_possible_places :192   (line numbers of the reference file)
"""
from __future__ import annotations

import numpy as np

from .... import boxes as _bx
from ....object_detection.Real3DAug.tools.find_spot import (DEFAULT, GREEN, RED, YELLOW, _correct_height, _probe,
                                                            _rotate_annotation, _transform_in_place)
from ....ops import cut_bounding_box_mask
from .cut_bbox import cut_bounding_box


def make_dictionary(annotation_array):
    return _bx.make_dictionary(annotation_array, ss=True)


def dictionary2array(annotation_dictionary):
    return _bx.dictionary2array(annotation_dictionary, ss=True)


def read_label_line(line):
    return _bx.read_label_line_ss(line)


def rotate_bounding_box_2(bbox_pcl, annotation, rotation=1):
    """Rotate the object's points (in place) and its annotation about the SENSOR z-axis (ss/fs:42-76)."""
    annotation, c, s = _rotate_annotation(annotation, rotation, ss=True)
    _transform_in_place(bbox_pcl, c, s, 0.0)
    return bbox_pcl, annotation


def check_bounding_box(scene_pcl, scene_anno, sample_pcl, sample_anno, ok_surface):
    """True iff no scene point with a label outside ``ok_surface`` lies in the candidate box and no object point lies
    in an existing box (ss/fs:79-104); one device call (r3d_obb_collide)."""
    from ....ops import obb_collide
    if len(ok_surface) > 8:                      # more surface labels than the primitive's label set holds
        inside = cut_bounding_box_mask(scene_pcl, sample_anno) & ~np.isin(scene_pcl[:, 7], ok_surface)
        return not inside.any() and not any(cut_bounding_box_mask(sample_pcl, anno).any() for anno in scene_anno)
    return not bool(obb_collide(scene_pcl, scene_anno, sample_pcl, [(1.0, 0.0, 0.0, sample_anno)], mode='ss',
                                ok_surface=[int(v) for v in ok_surface])[0])


def correct_height(scene_pcl, sample_pcl, sample_anno, ok_surface):
    """Road level under the box centre from the points labelled as an allowed surface (ss/fs:107-152)."""
    return _correct_height(scene_pcl, sample_pcl, sample_anno, list(ok_surface), ss=True)


def find_possible_places(point_cloud, scene_annotation, sample_data, map, map_move, original_pcl, transformation_matrix,
                         config):
    """All feasible yaw placements of a cut object on the (already occupancy-adjusted) rich map (ss/fs:192-273)."""
    sample_pcl = sample_data['pcl']
    anno_str = sample_data['anno']
    cls = int(read_label_line(anno_str.item())['class'][0])
    flags, boxes, xyz, rots, obj_anno = _probe('ss', point_cloud, scene_annotation, sample_pcl, anno_str, cls,
                                                original_pcl, config, map_data={'map': map, 'move': map_move},
                                                pose=np.asarray(transformation_matrix, dtype=np.float64))
    output_pcl, output_annotation, output_rotation = [], [], []
    for i, k in enumerate(rots):
        pcl = np.array(sample_pcl, copy=True)
        pcl[:, :3] = xyz[i]
        rec = (*boxes[k], obj_anno['length'], obj_anno['width'], obj_anno['height'])
        output_pcl.append(pcl)
        output_annotation.append(_bx.placed_box_dictionary(rec, str(cls), ss=True))
        output_rotation.append(k)
    not_on_road = int(np.sum((flags[1:] & 1) == 0))
    object_collision = int(np.sum(((flags[1:] & 3) == 3) & ((flags[1:] & 4) != 0)))
    print(f'From 360 possibilities, {YELLOW}{not_on_road}{DEFAULT} was not on road, {RED}{object_collision}{DEFAULT} has '
          f'collision with another object, and {GREEN}{len(output_pcl)}{DEFAULT} was possible.')
    return output_pcl, output_annotation, output_rotation
