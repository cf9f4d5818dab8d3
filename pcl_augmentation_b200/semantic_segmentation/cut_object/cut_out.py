"""Cut-object database of the semantic-segmentation pipeline on the GPU — the reference's
``semantic_segmentation/cut_object/cut_out.py`` (``ss/co``): for every annotated box of an insertable class, the
points strictly inside the box that carry the class label are saved as
``<bbox_path>/<label name>/<shortcut><sequence>-<frame>_<nn>_<ddd>_m.npz`` ``{anno, pcl: M x 5 float64}`` if they are at
least ``min_points`` (ss/co:96-157).

    python -m pcl_augmentation_b200.semantic_segmentation.cut_object.cut_out --sequence 00 [--config ...]
"""
import os

import numpy as np

from ... import boxes as bx
from ... import cut_objects as co


def line_box(annotation):
    """ss/co:120-139: box dictionary of one ``cls x y z h l w yaw`` line."""
    items = annotation.split(' ')
    x, y, z = float(items[1]), float(items[2]), float(items[3])
    height, width, length = float(items[4]), float(items[6]), float(items[5])
    q = bx._yaw_quaternion(float(items[7]))
    return bx.make_dictionary([[x, y, z], [q[0], q[1], q[2], q[3]], [width, length, height], [int(items[0])]], ss=True), x, y


def cut_frames(frames, config):
    """``frames``: list of (xyzi float32 N x 4, labels N, annotation lines, sequence, frame name).  Returns per frame the
    list of (label folder, file name, annotation line, pcl M x 5 float64) the reference would save."""
    classes = config['insertion']['classes']
    parsed, boxes, emit = [], [], []
    for xyzi, labels, lines, sequence, name in frames:
        rows = []
        for annotation in lines:
            if len(annotation) == 0:
                break
            cls = int(annotation.split(' ')[0])
            if cls not in classes:
                continue
            box, x, y = line_box(annotation)
            rows.append((cls, annotation, box, x, y))
        parsed.append(rows)
        boxes.append([r[2] for r in rows])
        emit.append([r[0] for r in rows])                                                    # ss/co:143
    if not any(boxes):
        return [[] for _ in frames]
    cuts = co.cut_boxes_batch([(f[0], f[1]) for f in frames], boxes, emit)
    out = []
    for f, rows in enumerate(parsed):
        sequence, name = frames[f][3], frames[f][4]
        classes_count = np.zeros(len(classes))
        saved = []
        for j, (cls, annotation, box, x, y) in enumerate(rows):
            classes_count[classes.index(cls)] += 1                                           # ss/co:118
            cut = cuts[f][j]
            if len(cut.xyzi) < config['insertion']['min_points'][cls]:
                continue
            pcl = np.hstack((cut.xyzi.astype(np.float64), cut.labels.astype(np.float64).reshape(-1, 1)))
            shortcut = config['insertion']['labels_shortcut'][cls]
            fname = f'{shortcut}{sequence}-{name}_{int(classes_count[classes.index(cls)]):02d}_{int(np.sqrt(x ** 2 + y ** 2)):03d}_m'
            saved.append((config['labels'][cls], fname, annotation, pcl))
        out.append(saved)
    return out


def generate_samples(config, sequence, batch_size=32, log=print):
    """The reference script's loop (ss/co:69-157) over one SemanticKITTI sequence."""
    from ..Real3DAug.tools.datasets import SemanticKITTI
    dataset = SemanticKITTI(config, sequence)
    save_path = config['path']['bbox_path']
    os.makedirs(save_path, exist_ok=True)
    for c in config['insertion']['classes']:
        os.makedirs(f'{save_path}/{config["labels"][c]}', exist_ok=True)
    n, saved = len(dataset), 0
    for i0 in range(0, n, batch_size):
        frames = []
        for i in range(i0, min(i0 + batch_size, n)):
            xyzi, labels, _, anno_file, name = dataset.read_frame(i)
            if not os.path.exists(anno_file):                                                # ss/co:100-101
                continue
            with open(anno_file, 'r') as f:
                lines = f.readlines()
            frames.append((xyzi, labels, lines, sequence, name))
        if frames:
            for per_frame in cut_frames(frames, config):
                for folder, fname, annotation, pcl in per_frame:
                    np.savez(f'{save_path}/{folder}/{fname}', anno=annotation, pcl=pcl)
                    saved += 1
        log(f'{min(i0 + batch_size, n)} / {n} frames, {saved} samples')
    return saved


def main(argv=None):
    import argparse
    import yaml
    ap = argparse.ArgumentParser(description="cut-object database (SemanticKITTI) on the GPU")
    ap.add_argument("--config", default="../config/semantic-kitti.yaml")
    ap.add_argument("--sequence", required=True)
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args(argv)
    with open(args.config, "r") as f:
        generate_samples(yaml.safe_load(f), str(args.sequence).zfill(2), args.batch)


if __name__ == "__main__":
    main()
