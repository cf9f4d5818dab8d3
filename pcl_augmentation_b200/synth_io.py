"""Write a ``synth.Case`` as an on-disk dataset in the reference's directory layouts and file formats, so that the
reference's ``insertion.py`` (oracle/make_golden.py) and this package's dataset drivers read the very same files.

OD (KITTI, od/ds:40-71, od/ins:303-362): ``data/velodyne/<f>.bin`` float32 N x 4, ``labels/<f>.label`` uint32,
``data/label_2/<f>.txt``, ``maps/maps/{road_maps,pedestrian_area}/npz/<f>.npz``, ``samples/<class>/<name>.npz``,
``train.txt``, ``config/KITTI.yaml``.
Semseg (SemanticKITTI, ss/ds:20-70, ss/ins:290-325): ``data/sequences/<s>/{velodyne,labels}``, ``poses.txt``,
``anno/sequences/<s>/bbox/<f>.txt``, ``maps/<s>.npz``, ``samples/<label name>/<name>.npz``.
"""
from __future__ import annotations

import json
import os

import numpy as np
import yaml

VELO_2_CAM = np.array([[7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03],
                       [1.480249e-02, 7.280733e-04, -9.998902e-01, -7.631618e-02],
                       [9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01], [0, 0, 0, 1]])      # ss/ds:22-25
MY_CALIB = np.array([[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 0], [0, 0, 0, 1.0]])                     # ss/ds:26-29


def write_od_dataset(cases, root, fixed_counts=None):
    """``cases``: list of OD cases (frame i -> ``%06d`` % i); they share the cut-object DB of the first one.
    Returns (cwd to run ``insertion.py`` from, output folder ``out/<chosen|random>/00``, config dict)."""
    d = lambda *p: os.path.join(root, *p)
    for p in ("config", "run", "data/velodyne", "data/label_2", "data/calib", "data/image_2", "labels", "out",
              "maps/maps/pedestrian_area/npz", "maps/maps/road_maps/npz"):
        os.makedirs(d(p))
    for i, case in enumerate(cases):
        name = f"{i:06d}"
        case.pcl5[:, :4].astype(np.float32).tofile(d(f"data/velodyne/{name}.bin"))
        case.pcl5[:, 4].astype(np.uint32).tofile(d(f"labels/{name}.label"))
        with open(d(f"data/label_2/{name}.txt"), "w") as f:
            for line in case.box_lines:
                f.write(line + "\n")
        np.savez(d(f"maps/maps/road_maps/npz/{name}.npz"), **case.maps["Road"])
        np.savez(d(f"maps/maps/pedestrian_area/npz/{name}.npz"), **case.maps["Sidewalk"])
    with open(d("train.txt"), "w") as f:
        for i in range(len(cases)):
            f.write(f"{i}\n")
    case = cases[0]
    for cls, items in case.db.items():
        os.makedirs(d("samples", str(cls)))
        for name, s in items:
            np.savez(d("samples", str(cls), name + ".npz"), pcl=s["pcl"], anno=s["anno"])
    cfg = json.loads(json.dumps(case.config))
    cfg["path"] = dict(dataset_path=d("data"), maps_path=d("maps"), label_path=d("labels"),
                       sample_path=d("samples"), output_path=d("out"), train_txt_path=d("train.txt"))
    if fixed_counts is not None:
        cfg["insertion"]["random"] = False
        cfg["insertion"]["number_of_classes"] = [int(c) for c in fixed_counts]
    with open(d("config/KITTI.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)
    folder = "random" if cfg["insertion"]["random"] else "chosen"
    return d("run"), d(f"out/{folder}/00"), cfg


def pose_file_row(pose):
    """KITTI odometry pose row (3 x 4, camera frame) whose ``create_transform_matrix`` (ss/ds:65-70) gives ``pose``."""
    return (MY_CALIB @ pose @ np.linalg.inv(VELO_2_CAM))[:3].reshape(1, 12)


def write_ss_dataset(cases, root, sequence="00", fixed_counts=None):
    """``cases``: list of semseg cases of ONE sequence (they share map and DB).  Returns (cwd, output folder
    ``out/<chosen|random>/00/sequences/<sequence>``, config dict)."""
    d = lambda *p: os.path.join(root, *p)
    seq = f"data/sequences/{sequence}"
    for p in ("config", "run", f"{seq}/velodyne", f"{seq}/labels", f"anno/sequences/{sequence}/bbox", "maps", "out",
              "samples"):
        os.makedirs(d(p))
    rows = []
    for i, case in enumerate(cases):
        name = f"{i:06d}"
        case.pcl5[:, :4].astype(np.float32).tofile(d(f"{seq}/velodyne/{name}.bin"))
        case.pcl5[:, 4].astype(np.uint32).tofile(d(f"{seq}/labels/{name}.label"))
        with open(d(f"anno/sequences/{sequence}/bbox/{name}.txt"), "w") as f:
            for line in case.box_lines:
                f.write(line + "\n")
        rows.append(pose_file_row(case.pose))
    if len(rows) < 2:
        rows = rows * 2                                           # >= 2 rows: np.loadtxt stays 2-D
    np.savetxt(d(f"{seq}/poses.txt"), np.concatenate(rows, axis=0), fmt="%.17g")
    case = cases[0]
    np.savez(d(f"maps/{sequence}.npz"), **case.map_data)
    for cls, items in case.db.items():
        folder = case.config["labels"][cls]
        os.makedirs(d("samples", folder))
        for name, s in items:
            np.savez(d("samples", folder, name + ".npz"), pcl=s["pcl"], anno=s["anno"])
    cfg = yaml.safe_load(yaml.safe_dump(case.config))            # deep copy that keeps the int keys
    cfg["path"] = dict(dataset_path=d("data"), maps_path=d("maps"), annotation_path=d("anno"),
                       bbox_path=d("samples"), output_path=d("out"))
    if fixed_counts is not None:
        cfg["insertion"]["random"] = False
        cfg["insertion"]["number_of_classes"] = [int(c) for c in fixed_counts]
    with open(d("config/semantic-kitti.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)
    folder = "random" if cfg["insertion"]["random"] else "chosen"
    return d("run"), d(f"out/{folder}/00/sequences/{sequence}"), cfg
