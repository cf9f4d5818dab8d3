"""Drop-in for the function layer of the reference's semantic-segmentation ``insertion.py`` (projection functions
:54-129, ``addjust_map_2`` :202, ``generate_seed`` :171); CUDA inside.  The script loop (:317-599) is replaced by
``pcl_augmentation_b200.dataset_driver.augment_semantic_kitti`` on top of ``Real3DEngine('ss', ...).augment_batch``:

    python -m pcl_augmentation_b200.semantic_segmentation.Real3DAug.insertion --sequence 0 [--config ../config/semantic-kitti.yaml]
    python -m pcl_augmentation_b200.semantic_segmentation.Real3DAug.insertion --dataset waymo [--config ../config/waymo.yaml]
"""
import numpy as np

from ... import _lib
from ... import ops as _ops
from ...object_detection.Real3DAug.insertion import generate_seed                   # noqa: F401
from ...ops import _dev, _stream, add_space_for_spherical, fill_spherical           # noqa: F401
from .tools.closing import class_closing, smooth_out                                # noqa: F401
from .tools.find_spot import *                                                      # noqa: F401,F403
from .tools.find_spot import read_label_line

NUMROW = 112
NUMCOLUMN = 360 * 4
MAX_NUM_TRIES = 100
ROAD_INDEXES = [40, 44, 48]     # undefined in the reference's semseg insertion.py (:209); value of od/fs:14


def geometrical_front_view(point_cloud, num_row, num_column, max_elevation_angle, min_elevation_angle, sample=False):
    _ops.NUMCOLUMN = NUMCOLUMN
    return _ops.geometrical_front_view(point_cloud, num_row, num_column, max_elevation_angle, min_elevation_angle, sample)


def extract_anno(anno_path):
    with open(anno_path, 'r') as f:
        return np.array([read_label_line(line) for line in f if len(line) > 0])


def addjust_map_2(map_data, point_cloud, transformation_matrix):
    """Mark the rich-map cells occupied by non-ground scene points below z < 1.5 with 4 (:202-224).  Returns
    (map, map_move) like the reference; cells outside the map are skipped (the reference would raise / wrap)."""
    lib = _lib.load()
    map_arr = map_data['map']
    map_move = map_data['move']
    d_map = _dev(map_arr, np.float64)
    if len(point_cloud):
        d_pts = _dev(point_cloud, np.float64)
        pose = np.ascontiguousarray(transformation_matrix, dtype=np.float64)
        ground = np.asarray(ROAD_INDEXES, dtype=np.int32)
        mv = np.asarray(map_move).reshape(-1)
        _lib.check(lib.r3d_adjust_map(d_pts.data_ptr(), len(point_cloud), pose.ctypes.data, int(mv[0]), int(mv[1]),
                                      ground.ctypes.data, len(ground), d_map.data_ptr(), map_arr.shape[0],
                                      map_arr.shape[1], _stream()), "addjust_map_2")
    out = d_map.cpu().numpy()
    if isinstance(map_arr, np.ndarray) and map_arr.flags.writeable:
        map_arr[...] = out                   # the reference edits the array it was handed in place
        out = map_arr
    return out, map_move


def main(argv=None):
    import argparse
    import yaml
    from ...dataset_driver import augment_semantic_kitti
    ap = argparse.ArgumentParser(description="Real3D-Aug insertion (semantic segmentation) on the CUDA engine")
    ap.add_argument("--dataset", default="semantic-kitti", choices=["semantic-kitti", "waymo"])        # ss/ins:227-243
    ap.add_argument("--config", default=None)
    ap.add_argument("--sequence", default=None, help="SemanticKITTI sequence (must be in split.train, ss/ins:254-258)")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--folder", type=int, default=None)
    ap.add_argument("--reverse", action="store_true", help="process the frames in reverse order (ss/ds:117-141)")
    ap.add_argument("--skip", type=int, default=0, help="skip the first frames of the sequence (ss/ds:117-141)")
    ap.add_argument("--yaw-steps", type=int, default=360)
    ap.add_argument("--processes", type=int, default=1,
                    help="start this many copies of the script on the same output folder (the reference's scale-out: the "
                         "frame markers of ss/ins:327-339 keep them apart); the GPU is shared")
    args = ap.parse_args(argv)
    if args.processes > 1:
        import subprocess
        import sys
        passed = [a for a in (sys.argv[1:] if argv is None else list(argv))]
        keep, skip = [], False
        for a in passed:                                     # the same command line without --processes N
            if skip:
                skip = False
                continue
            if a == "--processes":
                skip = True
                continue
            if a.startswith("--processes="):
                continue
            keep.append(a)
        if "--folder" not in keep:
            keep += ["--folder", "0"]
        mod = __spec__.name if __spec__ else "pcl_augmentation_b200.semantic_segmentation.Real3DAug.insertion"
        procs = [subprocess.Popen([sys.executable, "-m", mod] + keep) for _ in range(args.processes)]
        rc = max(p.wait() for p in procs)
        if rc:
            raise SystemExit(rc)
        return
    cfg_path = args.config or ("../config/semantic-kitti.yaml" if args.dataset == "semantic-kitti" else "../config/waymo.yaml")
    with open(cfg_path, "r") as f:
        config = yaml.safe_load(f)
    if args.dataset == "semantic-kitti":
        assert args.sequence is not None and int(args.sequence) in config["split"]["train"], "choose a training sequence"
        runs = [(f"{int(args.sequence):02d}", None)]
    else:
        from .tools.datasets import Waymo
        probe = Waymo(config)
        runs = []
        for seq in probe.sequence_names:
            ds = Waymo(config, reverse=args.reverse, skip_scenes=0)
            ds.velodyne_list = ds.velodyne_list[[f.split("/")[-3] == seq for f in ds.velodyne_list]]
            ds.sequence = seq
            runs.append((seq, ds))
    for seq, ds in runs:
        folder, written, skipped = augment_semantic_kitti(config, seq, batch_size=args.batch, yaw_steps=args.yaw_steps,
                                                          folder_number=args.folder, reverse=args.reverse,
                                                          skip_scenes=args.skip, dataset=ds)
        print(f"{folder}: {written} frames written, {skipped} without an insertion")


if __name__ == "__main__":
    main()
