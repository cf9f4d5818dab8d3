"""Drop-in for the reference's ``tools/cut_bbox.py`` (cut_bounding_box :7, separate_bbox :71), computed in CUDA."""
from ....ops import cut_bounding_box, cut_bounding_box_mask


def separate_bbox(point_cloud, annotation, annotation_move=[0, 0, 0]):
    """(scene without the box, points of the box).  Imported but never called by the reference (ss/ins:14).  Known
    deviation: the reference's version keeps points lying exactly ON a face in the box (its six tests are the
    non-strict complements, cb:87-110); here they stay in the scene, like ``cut_bounding_box``."""
    inside = cut_bounding_box_mask(point_cloud, annotation, annotation_move)
    return point_cloud[~inside], point_cloud[inside]


__all__ = ["cut_bounding_box", "separate_bbox"]
