"""Drop-in for the function layer of the reference's object-detection ``insertion.py``.

The projection functions (add_space_for_spherical :55, fill_spherical :68, geometrical_front_view :85) run in CUDA;
``generate_seed`` (:171), ``extract_anno`` (:133) and ``create_annotation_line`` (:227) are the reference's tiny host
helpers.  The script part of the reference (its ``while len(dataset_functions) > 0`` loop, :321-628) is
``pcl_augmentation_b200.dataset_driver.augment_kitti`` on top of ``Real3DEngine.augment_batch``:

    python -m pcl_augmentation_b200.object_detection.Real3DAug.insertion [--config ../config/KITTI.yaml] [--batch 64]

reads the reference's YAML (:293-294), walks ``train.txt`` and writes ``<output_path>/<random|chosen>/<NN>/``
``{velodyne, check, label_2, added_objects}`` like the reference.
"""
import numpy as np

from ... import ops as _ops
from ...boxes import create_annotation_line                      # noqa: F401
from ...ops import add_space_for_spherical, fill_spherical       # noqa: F401
from .tools.closing import class_closing, smooth_out             # noqa: F401
from .tools.find_spot import *                                   # noqa: F401,F403
from .tools.find_spot import read_label_line

NUMROW = 112
NUMCOLUMN = 360 * 4
MAX_NUM_TRIES = 100


def geometrical_front_view(point_cloud, num_row, num_column, max_elevation_angle, min_elevation_angle, sample=False):
    _ops.NUMCOLUMN = NUMCOLUMN          # pix_id uses this module's global, like the reference (:117)
    return _ops.geometrical_front_view(point_cloud, num_row, num_column, max_elevation_angle, min_elevation_angle, sample)


def extract_anno(anno_path):
    """List of box dictionaries from a KITTI label_2 file (:133-157)."""
    with open(anno_path, 'r') as f:
        return np.array([read_label_line(line) for line in f if len(line) > 0])


def generate_seed(config):
    """Objects to insert per class and the first class to insert (:171-187)."""
    if config['insertion']['random']:
        seed = np.zeros(len(config['insertion']['classes']))
        for i in np.random.randint(len(config['insertion']['classes']), size=config['insertion']['number_of_object']):
            seed[i] += 1
    else:
        seed = np.array(config['insertion']['number_of_classes'])
    inserted_class = None
    for i in range(len(seed)):
        if seed[i] > 0:
            inserted_class = config['insertion']['classes'][i]
            break
    return seed, inserted_class


def main(argv=None):
    import argparse
    import yaml
    from ...dataset_driver import augment_kitti
    ap = argparse.ArgumentParser(description="Real3D-Aug insertion (KITTI object detection) on the CUDA engine")
    ap.add_argument("--config", default="../config/KITTI.yaml")          # od/ins:293
    ap.add_argument("--batch", type=int, default=64, help="frames augmented per engine batch")
    ap.add_argument("--folder", type=int, default=None, help="run folder number 0-99 (default 00, od/ds:130)")
    ap.add_argument("--yaw-steps", type=int, default=360, help="yaw candidates per cut object (reference: 360)")
    ap.add_argument("--processes", type=int, default=1,
                    help="start this many copies of the script on the same output folder (the reference's own scale-out, "
                         "object_detection/README.md:36: the frame markers of od/ins:335-347 keep them apart); the host "
                         "side of a copy (file I/O, schedule draws) is one Python process, the GPU is shared")
    args = ap.parse_args(argv)
    if args.processes > 1:
        import subprocess
        import sys
        cmd = [sys.executable, "-m", __spec__.name if __spec__ else "pcl_augmentation_b200.object_detection.Real3DAug.insertion",
               "--config", args.config, "--batch", str(args.batch), "--yaw-steps", str(args.yaw_steps), "--folder",
               str(0 if args.folder is None else args.folder)]
        procs = [subprocess.Popen(cmd) for _ in range(args.processes)]
        rc = max(p.wait() for p in procs)
        if rc:
            raise SystemExit(rc)
        return
    with open(args.config, "r") as f:
        config = yaml.safe_load(f)
    folder, written, skipped = augment_kitti(config, batch_size=args.batch, yaw_steps=args.yaw_steps,
                                             folder_number=args.folder)
    print(f"{folder}: {written} frames written, {skipped} without an insertion")


if __name__ == "__main__":
    main()
