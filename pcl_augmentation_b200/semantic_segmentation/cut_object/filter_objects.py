"""Drop-in for the reference's ``semantic_segmentation/cut_object/filter_objects.py`` (``ss/fo``): per class and 1 m
distance bucket, a sample with fewer points than the average of the samples sharing its 1-degree yaw bin is deleted
(ss/fo:88-115).  Pure bookkeeping on file names, annotations and point counts — there is no point arithmetic to move
to the GPU; kept so the database workflow (cut_out -> filter_objects -> insertion) is complete.

    python -m pcl_augmentation_b200.semantic_segmentation.cut_object.filter_objects [--config ...]
"""
import glob
import os

import numpy as np


def rotation_bin(annotation):
    return int(np.rad2deg(float(str(annotation).split(' ')[7])) + 180)                       # ss/fo:101


def to_delete(samples):
    """``samples``: list of (name, annotation, number of points) of ONE class and ONE distance bucket."""
    total, number = np.zeros(360), np.zeros(360)
    for _, anno, n in samples:
        total[rotation_bin(anno)] += n
        number[rotation_bin(anno)] += 1
    with np.errstate(divide='ignore', invalid='ignore'):
        avg = np.where(number != 0, total / number, np.inf)                                  # ss/fo:106
    return [name for name, anno, n in samples if avg[rotation_bin(anno)] > n]               # ss/fo:114


def filter_samples(config, log=print):
    save_path = config['path']['bbox_path']
    classes = config['insertion']['classes']
    assert os.path.exists(f'{save_path}'), 'Root folder does not exist'
    for c in classes:
        assert os.path.exists(f'{save_path}/{config["labels"][c]}'), f'{config["labels"][c]} folder does not exist'
    removed = 0
    for c in classes:
        cl = config['labels'][c]
        for i in range(100):
            files = sorted(glob.glob(f'{save_path}/{cl}/*_{i:03d}_m.npz'))
            if not files:
                continue
            samples = []
            for f in files:
                z = np.load(f, allow_pickle=True)
                samples.append((f, z['anno'], len(z['pcl'])))
            for f in to_delete(samples):
                os.remove(f)
                removed += 1
        log(f'{cl}: {removed} samples removed so far')
    return removed


def main(argv=None):
    import argparse
    import yaml
    ap = argparse.ArgumentParser(description="remove cut objects sparser than their yaw bin's average")
    ap.add_argument("--config", default="../config/semantic-kitti.yaml")
    args = ap.parse_args(argv)
    with open(args.config, "r") as f:
        filter_samples(yaml.safe_load(f))


if __name__ == "__main__":
    main()
