"""Cut-object database of the object-detection pipeline on the GPU — the reference's
``object_detection/cut_object/object_cut_out.py`` (``od/co``): for every un-occluded ``label_2`` box of an insertable
class, the points of the box grown by 0.2 m must all lie in the camera's field of view (od/co:137-145, cutout.py); the
points of the box itself, without the ground labels, are saved as ``<sample_path>/<class>/<shortcut><frame>_<n>_<d>_m.npz``
``{anno, pcl: M x 5 float64 (x, y, z, intensity, 1)}`` if they are at least ``min_points`` (od/co:147-168).

    python -m pcl_augmentation_b200.object_detection.cut_object.object_cut_out [--config ../config/KITTI.yaml]
"""
import copy
import math
import os

import numpy as np

from ... import boxes as bx
from ... import cut_objects as co


def label_boxes(annotation):
    """od/co:93-139: (class, occluded, base box, expanded box, corrected x, corrected y) of one ``label_2`` line."""
    items = annotation.split(' ')
    h, w, l = float(items[8]), float(items[9]), float(items[10])
    x, y, z = float(items[11]), float(items[12]), float(items[13])
    corrected_x, corrected_y, corrected_z = float(z) + 0.27, float(x) * -1, float(y) * -1 - 0.08
    q = bx._yaw_quaternion(float(items[14]) * -1)
    base = bx.make_dictionary([[corrected_x, corrected_y, corrected_z], [q[0], q[1], q[2], q[3]], [w + 0.2, l + 0.2, h + 0.1],
                               [items[0]]])
    expand = copy.deepcopy(base)
    expand['length'] += 0.2
    expand['width'] += 0.2
    expand['height'] += 0.2
    return items[0], int(items[2]), base, expand, corrected_x, corrected_y


def image_shape(img_file):
    """(rows, cols) of the camera image (cutout.py:115-116) — read from the file header, the pixels are not decoded."""
    from PIL import Image
    with Image.open(img_file) as im:
        return im.size[1], im.size[0]


def cut_frames(frames, config):
    """``frames``: list of (xyzi float32 N x 4, labels N, label_2 lines, calib dict, image shape, frame name).
    Returns per frame the list of (class, file name, annotation line, pcl M x 5 float64) the reference would save."""
    classes = config['insertion']['classes']
    drop = [config['labels'][k] for k in ('Road', 'Parking', 'Sidewalk')]                    # od/co:150-152
    parsed, boxes, emit, use_drop, cams = [], [], [], [], []
    for xyzi, labels, lines, calib, img_shape, name in frames:
        rows = []
        for annotation in lines:
            if len(annotation) == 0:
                break
            if annotation.split(' ')[0] not in classes:
                continue
            cls, occluded, base, expand, cx, cy = label_boxes(annotation)
            if occluded != 0:                                                                # od/co:104-107
                continue
            rows.append((cls, annotation, base, expand, cx, cy))
        parsed.append(rows)
        boxes.append([b for r in rows for b in (r[3], r[2])])                                # expanded, base
        emit.append([e for _ in rows for e in (co.EMIT_NONE, co.EMIT_ANY)])
        use_drop.append([u for _ in rows for u in (0, 1)])
        cams.append(co.camera_record(calib, img_shape))
    if not any(boxes):
        return [[] for _ in frames]
    cuts = co.cut_boxes_batch([(f[0], f[1]) for f in frames], boxes, emit, use_drop, drop, cams)
    out = []
    for f, rows in enumerate(parsed):
        name = frames[f][5]
        classes_count = np.zeros(len(classes))
        saved = []
        for j, (cls, annotation, base, expand, cx, cy) in enumerate(rows):
            grown, own = cuts[f][2 * j], cuts[f][2 * j + 1]
            if grown.count_inside != grown.count_fov:                                        # od/co:144
                continue
            classes_count[classes.index(cls)] += 1
            if len(own.xyzi) < config['insertion']['min_points'][cls]:                       # od/co:156
                continue
            pcl = np.hstack((own.xyzi.astype(np.float64), np.ones((len(own.xyzi), 1))))      # od/co:154-160
            fname = (f'{config["insertion"]["labels_shortcut"][cls]}{name}_{int(classes_count[classes.index(cls)])}_'
                     f'{int(np.sqrt(cx ** 2 + cy ** 2))}_m')
            saved.append((cls, fname, annotation, pcl))
        out.append(saved)
    return out


def generate_samples(config, batch_size=32, log=print):
    """The reference script's loop (od/co:61-168) over the ``train.txt`` frames, ``batch_size`` frames per launch."""
    from ..Real3DAug.tools.datasets import KITTI
    dataset = KITTI(config)
    save_path = config['path']['sample_path']
    for c in config['insertion']['classes']:
        os.makedirs(f'{save_path}/{c}', exist_ok=True)
    n, saved = len(dataset), 0
    for i0 in range(0, n, batch_size):
        frames = []
        for i in range(i0, min(i0 + batch_size, n)):
            xyzi, labels, name = dataset.read_frame(i)
            _, label_address, _, calib_file, img_file = dataset[i]
            with open(label_address, 'r') as f:
                lines = f.readlines()
            frames.append((xyzi, labels, lines, co.read_kitti_calib(calib_file), image_shape(img_file), name))
        for per_frame in cut_frames(frames, config):
            for cls, fname, annotation, pcl in per_frame:
                np.savez(f'{save_path}/{cls}/{fname}', anno=annotation, pcl=pcl)
                saved += 1
        log(f'{min(i0 + batch_size, n)} / {n} frames, {saved} samples')
    return saved


def main(argv=None):
    import argparse
    import yaml
    ap = argparse.ArgumentParser(description="cut-object database (KITTI) on the GPU")
    ap.add_argument("--config", default="../config/KITTI.yaml")
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args(argv)
    with open(args.config, "r") as f:
        generate_samples(yaml.safe_load(f), args.batch)


if __name__ == "__main__":
    main()
