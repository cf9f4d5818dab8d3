"""Shared helpers for the parity tests (test infrastructure)."""
import json
import math
import os

import numpy as np

from pcl_augmentation_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_SHAPE = synth.ScanShape(32, 600, 2.0, -24.8)
CASE_DEFAULTS = dict(shape=GOLDEN_SHAPE, n_per_class=100, obj_range=(4.0, 16.0))


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def case_from_spec(spec):
    kw = dict(CASE_DEFAULTS)
    kw.update({k: v for k, v in spec.items() if k not in ("task", "seed")})
    if isinstance(kw.get("shape"), (list, tuple)):
        kw["shape"] = synth.ScanShape(*kw["shape"])
    if "obj_range" in kw:
        kw["obj_range"] = tuple(kw["obj_range"])
    return synth.make_case(spec["task"], spec["seed"], **kw)


def case_from_golden(g):
    spec = json.loads(str(g["meta"]))["spec"]
    case = case_from_spec(spec)
    assert synth.case_digest(case) == str(g["digest"]), "regenerated synthetic inputs differ from the golden run"
    return spec, case


def parse_inserted(text):
    """'<name> with rotation: <rot>' lines of added_objects/<frame>.txt (od/ins:537)."""
    out = []
    for line in text.splitlines():
        if line.strip():
            name, rot = line.split(" with rotation: ")
            out.append((name, int(rot)))
    return out


RICH_MAP_SS_SEEDS = (41, 42, 43, 44)


def rich_map_ss_cases():
    """Four frames of one synthetic sequence: the poses are chained so that the frames overlap (cells written by
    several frames and by several surface classes: the last-writer and the sticky-sidewalk rules both matter)."""
    cases = [case_from_spec(dict(task="ss", seed=s, counts=[1] * 8, n_cars=c)) for s, c in zip(RICH_MAP_SS_SEEDS, (0, 3, 5, 2))]
    base = cases[0].pose
    for i, case in enumerate(cases[1:], 1):
        a = 0.35 * i
        step = np.eye(4)
        step[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        step[:3, 3] = [7.3 * i, -2.1 * i, 0.05 * i]
        case.pose = base @ step
    return cases


# ------------------------------------------------------------------------------------ cut-object DB
KITTI_CALIB_LINES = [
    "P0: 7.215377e+02 0.000000e+00 6.095593e+02 0.000000e+00 0.000000e+00 7.215377e+02 1.728540e+02 0.000000e+00 0.000000e+00 0.000000e+00 1.000000e+00 0.000000e+00",
    "P1: 7.215377e+02 0.000000e+00 6.095593e+02 -3.875744e+02 0.000000e+00 7.215377e+02 1.728540e+02 0.000000e+00 0.000000e+00 0.000000e+00 1.000000e+00 0.000000e+00",
    "P2: 7.215377e+02 0.000000e+00 6.095593e+02 4.485728e+01 0.000000e+00 7.215377e+02 1.728540e+02 2.163791e-01 0.000000e+00 0.000000e+00 1.000000e+00 2.745884e-03",
    "P3: 7.215377e+02 0.000000e+00 6.095593e+02 -3.395242e+02 0.000000e+00 7.215377e+02 1.728540e+02 2.199936e+00 0.000000e+00 0.000000e+00 1.000000e+00 2.729905e-03",
    "R0_rect: 9.999239e-01 9.837760e-03 -7.445048e-03 -9.869795e-03 9.999421e-01 -4.278459e-03 7.402527e-03 4.351614e-03 9.999631e-01",
    "Tr_velo_to_cam: 7.533745e-03 -9.999714e-01 -6.166020e-04 -4.069766e-03 1.480249e-02 7.280733e-04 -9.998902e-01 -7.631618e-02 9.998621e-01 7.523790e-03 1.480755e-02 -2.717806e-01",
    "Tr_imu_to_velo: 9.999976e-01 7.553071e-04 -2.035826e-03 -8.086759e-01 -7.854027e-04 9.998898e-01 -1.482298e-02 3.195559e-01 2.024406e-03 1.482454e-02 9.998881e-01 -7.997231e-01",
]
KITTI_IMAGE_SHAPE = (375, 1242)
CUT_OBJECT_SEEDS = {"od": (81, 82, 83), "ss": (91, 92, 93)}


def cut_object_cases(task, per_class=6, shape=None):
    """Frames that CONTAIN objects of the insertable classes: cut objects of the synthetic DB are put back into the
    scan at their annotated place (their points carry the class label / a non-ground label, a few carry another label
    so that the label filters of the cut-out scripts matter) and their annotation line is added to the frame's
    annotations; in OD every fifth line is marked occluded (skipped by object_cut_out.py:104-107)."""
    cases = []
    for fi, seed in enumerate(CUT_OBJECT_SEEDS[task]):
        spec = dict(task=task, seed=seed, counts=[1] * (2 if task == "od" else 8), n_cars=3)
        if shape is not None:
            spec["shape"] = shape
            spec["obj_range"] = (5.0, 35.0)
        case = case_from_spec(spec)
        extra, lines = [], []
        k = 0
        for cls in case.config["insertion"]["classes"]:
            items = [it[1] for it in case.db[cls]]
            if task == "od":
                # two thirds of the picks lie in front of the car (inside the camera's field of view), the rest anywhere
                def azimuth(it):
                    f = str(it["anno"]).split(" ")
                    return abs(math.atan2(-float(f[11]), float(f[13]) + 0.27))
                front = [it for it in items if azimuth(it) < 0.55]
                picks = front[fi * 4:(fi + 1) * 4] + items[fi * 2:(fi + 1) * 2]
            else:
                picks = items[fi * per_class:(fi + 1) * per_class]
                if fi > 0:                     # the same objects again with half of their points (filter_objects.py)
                    picks = picks + [dict(pcl=it["pcl"][::2], anno=it["anno"]) for it in items[:2]]
            for s in picks:
                pts = s["pcl"].copy()
                line = str(s["anno"])
                if task == "od":
                    pts[:, 4] = 30.0
                    pts[::7, 4] = float(case.config["labels"]["Road"])
                    if k % 5 == 4:
                        items = line.split(" ")
                        items[2] = "1"
                        line = " ".join(items)
                else:
                    pts[:, 4] = float(cls)
                    pts[::9, 4] = 70.0
                extra.append(pts)
                lines.append(line)
                k += 1
        case.pcl5 = np.vstack([case.pcl5] + extra)
        case.box_lines = list(case.box_lines) + lines
        cases.append(case)
    return cases


def write_kitti_camera_files(root, n_frames):
    """calib/<f>.txt and image_2/<f>.png (black, KITTI size) next to the velodyne files of ``write_od_dataset``."""
    from PIL import Image
    for i in range(n_frames):
        with open(os.path.join(root, f"data/calib/{i:06d}.txt"), "w") as f:
            f.write("\n".join(KITTI_CALIB_LINES) + "\n")
        Image.new("RGB", (KITTI_IMAGE_SHAPE[1], KITTI_IMAGE_SHAPE[0])).save(os.path.join(root, f"data/image_2/{i:06d}.png"))


def read_sample_dir(folder):
    """{file name: (annotation str, pcl)} of a directory of cut-object ``.npz`` files."""
    out = {}
    for f in sorted(os.listdir(folder)):
        if f.endswith(".npz"):
            z = np.load(os.path.join(folder, f), allow_pickle=True)
            out[f[:-4]] = (str(z["anno"]), z["pcl"])
    return out
