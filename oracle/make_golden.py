"""TEST INFRASTRUCTURE ONLY — generates ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py [--only NAME]

Two kinds of fixture:
  * function-level: the reference's own functions (``fill_spherical``, ``geometrical_front_view``, ``smooth_out``,
    ``cut_bounding_box``, ``find_possible_places``, ``addjust_map_2``) called through ``oracle/shim.py`` on seeded
    synthetic inputs;
  * end-to-end: the reference's ``insertion.py`` executed as ``__main__`` (``runpy``) on a temporary on-disk
    dataset in the reference's formats; what it wrote (velodyne/check/label files, ``added_objects/*.txt``) is
    recorded.
Inputs are NOT stored: every fixture carries the ``make_case`` arguments plus a sha256 of the regenerated inputs, so
tests rebuild them from the seed and check the digest.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import shim                                        # noqa: E402
from pcl_augmentation_b200 import synth                        # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_SHAPE = synth.ScanShape(32, 600, 2.0, -24.8)            # 19 200 pts: the reference finishes in ~a minute

E2E_CASES = {
    "e2e_od_a": dict(task="od", seed=11, counts=[2, 2], n_cars=5),
    "e2e_od_b": dict(task="od", seed=12, counts=[1, 3], n_cars=9),
    "e2e_ss_a": dict(task="ss", seed=21, counts=[1, 0, 0, 2, 1, 0], n_cars=4),
    "e2e_ss_b": dict(task="ss", seed=22, counts=[0, 1, 1, 1, 0, 1], n_cars=7),
    # crowded scenes (many near misses in the collision tests) and more objects per scan
    "e2e_od_c": dict(task="od", seed=13, counts=[3, 3], n_cars=14),
    "e2e_ss_c": dict(task="ss", seed=23, counts=[1, 1, 1, 1, 1, 1], n_cars=13),
    # BASELINE-size scans (configs C1 / C2: 120 000 and 124 992 points, 112 x 1440 image), minutes per scan in the reference
    "e2e_od_full": dict(task="od", seed=14, counts=[2, 2], n_cars=6, shape=[64, 1875, 2.0, -24.8], obj_range=[5.0, 35.0]),
    "e2e_ss_full": dict(task="ss", seed=24, counts=[1, 1, 0, 1, 1, 0], n_cars=6, shape=[64, 1953, 2.0, -24.8],
                        obj_range=[5.0, 35.0]),
    # one class of very close, very large cut objects (> 4096 points, pixel rectangle > 8192 px): the global-scratch
    # branches of the candidate selection (r3d_k_occlusion.cuh)
    "e2e_ss_big": dict(task="ss", seed=25, counts=[1], n_cars=2, shape=[48, 1200, 2.0, -24.8], obj_range=[3.2, 4.2],
                       classes=[18], n_per_class=100),
}
CASE_DEFAULTS = dict(shape=GOLDEN_SHAPE, n_per_class=100, obj_range=(4.0, 16.0))


def build_case(spec):
    kw = dict(CASE_DEFAULTS)
    kw.update({k: v for k, v in spec.items() if k not in ("task", "seed")})
    if isinstance(kw.get("shape"), (list, tuple)):
        kw["shape"] = synth.ScanShape(*kw["shape"])
    if "obj_range" in kw:
        kw["obj_range"] = tuple(kw["obj_range"])
    return synth.make_case(spec["task"], spec["seed"], **kw)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def subsequence_mask(original_f32, out_f32):
    """keep mask of the original rows that appear (in order) at the head of ``out_f32``; returns (mask, n_matched)."""
    mask = np.zeros(len(original_f32), dtype=bool)
    j = 0
    for i in range(len(original_f32)):
        if j < len(out_f32) and np.array_equal(original_f32[i], out_f32[j]):
            mask[i] = True
            j += 1
    return mask, j


# ------------------------------------------------------------------------------------------------ e2e
def _write_od_dataset(case, root):
    d = lambda *p: os.path.join(root, *p)
    for p in ("config", "run", "data/velodyne", "data/label_2", "data/calib", "data/image_2", "labels", "out",
              "maps/maps/pedestrian_area/npz", "maps/maps/road_maps/npz"):
        os.makedirs(d(p))
    case.pcl5[:, :4].astype(np.float32).tofile(d("data/velodyne/000000.bin"))
    case.pcl5[:, 4].astype(np.uint32).tofile(d("labels/000000.label"))
    with open(d("data/label_2/000000.txt"), "w") as f:
        for l in case.box_lines:
            f.write(l + "\n")
    with open(d("train.txt"), "w") as f:
        f.write("0\n")
    np.savez(d("maps/maps/road_maps/npz/000000.npz"), **case.maps["Road"])
    np.savez(d("maps/maps/pedestrian_area/npz/000000.npz"), **case.maps["Sidewalk"])
    for cls, items in case.db.items():
        os.makedirs(d("samples", str(cls)))
        for name, s in items:
            np.savez(d("samples", str(cls), name + ".npz"), pcl=s["pcl"], anno=s["anno"])
    cfg = json.loads(json.dumps(case.config))
    cfg["path"] = dict(dataset_path=d("data"), maps_path=d("maps"), label_path=d("labels"),
                       sample_path=d("samples"), output_path=d("out"), train_txt_path=d("train.txt"))
    cfg["insertion"]["random"] = False
    cfg["insertion"]["number_of_classes"] = [int(c) for c in case.schedule.counts]
    with open(d("config/KITTI.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)
    return d("run"), d("out/chosen/00")


def _write_ss_dataset(case, root, ref):
    d = lambda *p: os.path.join(root, *p)
    for p in ("config", "run", "data/sequences/00/velodyne", "data/sequences/00/labels", "anno/sequences/00/bbox",
              "maps", "out", "samples"):
        os.makedirs(d(p))
    case.pcl5[:, :4].astype(np.float32).tofile(d("data/sequences/00/velodyne/000000.bin"))
    case.pcl5[:, 4].astype(np.uint32).tofile(d("data/sequences/00/labels/000000.label"))
    with open(d("anno/sequences/00/bbox/000000.txt"), "w") as f:
        for l in case.box_lines:
            f.write(l + "\n")
    # pose file such that the reference's create_transform_matrix (ss/ds:65-70) gives back (almost exactly) case.pose
    ds_cls = ref.datasets.SemanticKITTI
    probe = ds_cls.__new__(ds_cls)
    velo_2_cam = np.array([[7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03],
                           [1.480249e-02, 7.280733e-04, -9.998902e-01, -7.631618e-02],
                           [9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01], [0, 0, 0, 1]])
    my_calib = np.array([[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 0], [0, 0, 0, 1.0]])
    pose = my_calib @ case.pose @ np.linalg.inv(velo_2_cam)
    np.savetxt(d("data/sequences/00/poses.txt"), np.repeat(pose[:3].reshape(1, 12), 2, axis=0), fmt="%.17g")   # >= 2 rows: loadtxt stays 2-D
    probe.velo_2_cam, probe.my_calib = velo_2_cam, my_calib
    used_pose = ds_cls.create_transform_matrix(probe, np.loadtxt(d("data/sequences/00/poses.txt")), 0)
    np.savez(d("maps/00.npz"), **case.map_data)
    for cls, items in case.db.items():
        folder = case.config["labels"][cls]
        os.makedirs(d("samples", folder))
        for name, s in items:
            np.savez(d("samples", folder, name + ".npz"), pcl=s["pcl"], anno=s["anno"])
    cfg = yaml.safe_load(yaml.safe_dump(case.config))   # deep copy that keeps the int keys
    cfg["path"] = dict(dataset_path=d("data"), maps_path=d("maps"), annotation_path=d("anno"),
                       bbox_path=d("samples"), output_path=d("out"))
    cfg["insertion"]["random"] = False
    cfg["insertion"]["number_of_classes"] = [int(c) for c in case.schedule.counts]
    with open(d("config/semantic-kitti.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)
    return d("run"), d("out/chosen/00/sequences/00"), used_pose


def make_shuffle_fn(case, folder_to_class):
    state = {"event": 0}
    classes = case.config["insertion"]["classes"]

    def shuffle(lst):
        cls = folder_to_class[os.path.basename(os.path.dirname(lst[0]))]
        ci = classes.index(cls)
        ordered = sorted(lst)
        head = [int(v) for v in case.schedule.perms[state["event"]][ci] if v >= 0]
        hs = set(head)
        rest = [j for j in range(len(ordered)) if j not in hs]
        lst[:] = [ordered[j] for j in head + rest]
        state["event"] += 1

    return shuffle, state


def gen_e2e(name, spec):
    t0 = time.time()
    task = spec["task"]
    case = build_case(spec)
    ref = shim.load(task)
    root = tempfile.mkdtemp(prefix="r3d_golden_")
    try:
        used_pose = None
        if task == "od":
            cwd, out = _write_od_dataset(case, root)
            folder_to_class = {str(c): c for c in case.config["insertion"]["classes"]}
            inputs = []
        else:
            cwd, out, used_pose = _write_ss_dataset(case, root, ref)
            folder_to_class = {case.config["labels"][c]: c for c in case.config["insertion"]["classes"]}
            inputs = ["1", "0", "no"]
        shuffle_fn, state = make_shuffle_fn(case, folder_to_class)
        shim.run_main(task, cwd, inputs=inputs, shuffle_fn=shuffle_fn)
        fix = {"meta": json.dumps(dict(spec=spec, defaults=dict(shape=list(vars(GOLDEN_SHAPE).values()),
                                                                 n_per_class=100, obj_range=[4.0, 16.0]))),
               "digest": synth.case_digest(case), "n_events": state["event"]}
        added_txt = os.path.join(out, "added_objects", "000000.txt")
        if os.path.exists(added_txt):
            velodyne = np.fromfile(os.path.join(out, "velodyne", "000000.bin"), dtype=np.float32).reshape(-1, 4)
            check = np.fromfile(os.path.join(out, "check", "000000.bin"), dtype=np.float32)
            check = check.reshape(-1, 5 if task == "ss" else 4)
            mask, n_kept = subsequence_mask(case.pcl5[:, :4].astype(np.float32), velodyne)
            fix.update(inserted=open(added_txt).read(), velodyne_sha=sha(velodyne), n_out=len(velodyne),
                       keep_orig=np.packbits(mask), n_kept=n_kept, tail=velodyne[n_kept:], check=check)
            if task == "ss":
                labels = np.fromfile(os.path.join(out, "labels", "000000.label"), dtype=np.uint32)
                fix.update(labels_sha=sha(labels), tail_labels=labels[n_kept:])
            else:
                with open(os.path.join(out, "label_2", "000000.txt")) as f:
                    fix["label_2"] = f.read()
        else:
            fix.update(inserted="", n_out=-1)
        if used_pose is not None:
            fix["used_pose"] = used_pose
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **fix)
        print(f"{name}: {time.time() - t0:.1f}s inserted={fix['inserted']!r} n_out={fix['n_out']}")
    finally:
        shutil.rmtree(root, ignore_errors=True)


# --------------------------------------------------------------------------------------- function level
FN_IMG = (64, 512)


def gen_fn_projection():
    """A1-A4 on a small scan, image 64 x 512 (module globals NUMROW/NUMCOLUMN set accordingly)."""
    ref = shim.load("od")
    ref.set_image_size(*FN_IMG)
    try:
        pcl, labels = synth.make_scan(5, synth.ScanShape(32, 300, 2.0, -24.8))
        pcl5 = np.hstack((pcl, labels.reshape(-1, 1))).astype(np.float64)
        pc = ref.insertion.add_space_for_spherical(pcl5)
        pc, mx, mn = ref.insertion.fill_spherical(pc)
        train, label, pc = ref.insertion.geometrical_front_view(pc, FN_IMG[0], FN_IMG[1], mx, mn)
        s_train, s_label = ref.closing.smooth_out(train, label)
        # an "object" projected with the scene's elevation range and sample=True (od/ins:474-480)
        obj = synth.make_cut_object(77, "Cyclist", False, synth.ScanShape(32, 300, 2.0, -24.8), (6.0, 9.0))["pcl"]
        obj[:, 2] += np.linspace(-0.5, 3.0, len(obj))           # push some points outside the elevation range
        opc = ref.insertion.add_space_for_spherical(obj)
        opc, _, _ = ref.insertion.fill_spherical(opc)
        o_train, o_label, opc = ref.insertion.geometrical_front_view(opc, FN_IMG[0], FN_IMG[1], mx, mn, sample=True)
        os_train, os_label = ref.closing.smooth_out(o_train, o_label)
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, "fn_projection.npz"), in_digest=synth.array_digest(pcl5, obj),
            max_el=mx, min_el=mn, sph=pc[:, 3:6], pix=pc[:, 8].astype(np.int32), train=train,
            label=label.astype(np.int8), s_train=s_train, s_label=s_label.astype(np.int8),
            o_sph=opc[:, 3:6], o_pix=opc[:, 8].astype(np.int32), o_train=o_train, o_label=o_label.astype(np.int8),
            os_train=os_train, os_label=os_label.astype(np.int8))
        print("fn_projection: filled px", int(((s_label == 1) & (label != 1)).sum()),
              "obj skipped", int((opc[:, 8] < 0).sum()), "obj filled", int(((os_label == 1) & (o_label != 1)).sum()))
    finally:
        ref.set_image_size(112, 1440)


def gen_fn_projection_full():
    """A1-A4 at BASELINE size: a 120 000-point KITTI-shape scan on the reference's own 112 x 1440 image.  The images are
    1.3 MB each, so the fixture keeps the pix ids and the elevation range, and sha256 digests of the fp64 images (the
    bar is bit-exactness) plus the positions / values of the filled pixels for diagnosis."""
    ref = shim.load("od")
    ref.set_image_size(112, 1440)
    pcl, labels = synth.make_scan(6, synth.KITTI_SHAPE)
    pcl5 = np.hstack((pcl, labels.reshape(-1, 1))).astype(np.float64)
    t0 = time.time()
    pc = ref.insertion.add_space_for_spherical(pcl5)
    pc, mx, mn = ref.insertion.fill_spherical(pc)
    train, label, pc = ref.insertion.geometrical_front_view(pc, 112, 1440, mx, mn)
    s_train, s_label = ref.closing.smooth_out(train, label)
    filled = np.flatnonzero((s_label == 1) & (label != 1)).astype(np.int32)
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "fn_projection_full.npz"), in_digest=synth.array_digest(pcl5), max_el=mx, min_el=mn,
        pix=pc[:, 8].astype(np.int32), r_sha=sha(pc[:, 3]), train_sha=sha(train), label_sha=sha(label.astype(np.int8)),
        s_train_sha=sha(s_train), s_label_sha=sha(s_label.astype(np.int8)), filled=filled,
        filled_values=s_train.ravel()[filled], occupied=int((label == 1).sum()))
    print(f"fn_projection_full: {time.time() - t0:.1f}s occupied px {int((label == 1).sum())} filled px {len(filled)}")


def gen_fn_cut_bbox():
    ref = shim.load("od")
    rng = np.random.default_rng(99)
    pts = rng.uniform(-6, 6, (6000, 5))
    pts[:, 2] = rng.uniform(-2, 3, 6000)
    pts = pts.astype(np.float32).astype(np.float64)
    boxes, masks = [], []
    for i in range(12):
        q = rng.normal(size=4)
        if i < 6:
            q[:2] = 0                                              # yaw-only, like every box the pipelines build
        q /= np.linalg.norm(q)
        anno = {"center": {"x": rng.uniform(-2, 2), "y": rng.uniform(-2, 2), "z": rng.uniform(-1.5, 0)},
                "rotation": {"x": q[0], "y": q[1], "z": q[2], "w": q[3]},
                "length": rng.uniform(1, 6), "width": rng.uniform(1, 4), "height": rng.uniform(1, 3), "class": "Car"}
        inside = ref.cut_bbox.cut_bounding_box(pts, anno)
        m = np.zeros(len(pts), dtype=bool)
        # rows are unique, so membership identifies the mask
        lookup = {r.tobytes() for r in inside}
        for k in range(len(pts)):
            m[k] = pts[k].tobytes() in lookup
        assert m.sum() == len(inside)
        boxes.append([anno["center"]["x"], anno["center"]["y"], anno["center"]["z"], *q,
                      anno["length"], anno["width"], anno["height"]])
        masks.append(np.packbits(m))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "fn_cut_bbox.npz"), in_digest=synth.array_digest(pts),
                        boxes=np.array(boxes), masks=np.array(masks), counts=[int(np.unpackbits(m).sum()) for m in masks])
    print("fn_cut_bbox: inside counts", [int(np.unpackbits(m).sum()) for m in masks])


def _scene9(ref, case):
    pcl5 = case.pcl5.copy()
    if case.task == "od":
        pcl5[pcl5[:, 4] != case.config["labels"]["Road"], 4] = 1           # od/ins:353-355
    pc = ref.insertion.add_space_for_spherical(pcl5)
    pc, mx, mn = ref.insertion.fill_spherical(pc)
    _, _, pc = ref.insertion.geometrical_front_view(pc, 112, 1440, mx, mn)
    return pcl5, pc


def gen_fn_places(task):
    ref = shim.load(task)
    spec = dict(task=task, seed=31 if task == "od" else 41, counts=None, n_cars=6)
    case = build_case(spec)
    original, scene = _scene9(ref, case)
    classes = case.config["insertion"]["classes"]
    rec = {"meta": json.dumps(dict(spec=spec)), "digest": synth.case_digest(case)}
    if task == "od":
        annos = [ref.find_spot.read_label_line(l) for l in case.box_lines]
    else:
        annos = [ref.find_spot.read_label_line(l) for l in case.box_lines]
        md = shim.FreshDict(case.map_data)
        map_arr, map_move = ref.insertion.addjust_map_2(md, scene, case.pose)
        rec["map_adjusted_cells"] = np.argwhere(map_arr == 4).astype(np.int32)
    t0 = time.time()
    n_feasible = []
    for ci, cls in enumerate(classes):
        for j in range(4):                                            # 4 samples per class
            name, sample = case.db[cls][j]
            sd = shim.FreshDict(sample)
            if task == "od":
                placement = case.config["insertion"]["placement"][cls]
                out = ref.find_spot.find_possible_places(scene, annos, sd, shim.FreshDict(case.maps[placement]),
                                                         original, case.config)
            else:
                out = ref.find_spot.find_possible_places(scene, annos, sd, map_arr, map_move, original, case.pose,
                                                         case.config)
            pcls, ans, rots = out
            key = f"c{ci}_s{j}"
            rec[key + "_rots"] = np.array(rots, dtype=np.int32)
            n_feasible.append(len(rots))
            if rots:
                pick = sorted({0, len(rots) // 2, len(rots) - 1})
                rec[key + "_pick"] = np.array(pick, dtype=np.int32)
                rec[key + "_xyz"] = np.stack([pcls[p][:, :3] for p in pick])
                rec[key + "_box"] = np.array([[ans[p]["center"]["x"], ans[p]["center"]["y"], ans[p]["center"]["z"],
                                               ans[p]["rotation"]["x"], ans[p]["rotation"]["y"], ans[p]["rotation"]["z"],
                                               ans[p]["rotation"]["w"]] for p in pick])
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"fn_places_{task}.npz"), **rec)
    print(f"fn_places_{task}: {time.time() - t0:.1f}s feasible per sample {n_feasible}")


def gen_rich_map_od():
    """the reference's per-frame rich-map script on two synthetic frames (written with
    pcl_augmentation_b200.synth_io in the reference's dataset layout)."""
    from pcl_augmentation_b200 import synth_io
    t0 = time.time()
    cases = [build_case(dict(task="od", seed=s, counts=[1, 1], n_cars=c)) for s, c in ((31, 3), (32, 8))]
    root = tempfile.mkdtemp(prefix="r3d_golden_")
    try:
        cwd, _, cfg = synth_io.write_od_dataset(cases, root)
        shutil.rmtree(os.path.join(root, "maps"))                     # the script must create everything itself
        os.makedirs(os.path.join(root, "maps"))
        shim.run_rich_map_od(cwd)
        rec = {"meta": json.dumps(dict(seeds=[31, 32], n_cars=[3, 8])), "digest0": synth.case_digest(cases[0]),
               "digest1": synth.case_digest(cases[1])}
        for i in range(2):
            for key, sub in (("road", "road_maps"), ("ped", "pedestrian_area")):
                z = np.load(os.path.join(root, f"maps/maps/{sub}/npz/{i:06d}.npz"))
                m = z["map"]
                assert set(np.unique(m)) <= {0, 1}
                rec[f"{key}{i}_shape"] = np.array(m.shape)
                rec[f"{key}{i}_dtype"] = str(m.dtype)
                rec[f"{key}{i}_bits"] = np.packbits(m.astype(bool))
                rec[f"{key}{i}_min"] = np.array([int(z["min_x"]), int(z["min_y"])])
        np.savez_compressed(os.path.join(GOLDEN_DIR, "rich_map_od.npz"), **rec)
        print(f"rich_map_od: {time.time() - t0:.1f}s shapes", rec["road0_shape"], rec["road1_shape"], rec["road0_dtype"],
              "road cells", int(np.unpackbits(rec["road0_bits"]).sum()), "ped cells", int(np.unpackbits(rec["ped0_bits"]).sum()))
    finally:
        shutil.rmtree(root, ignore_errors=True)


def gen_rich_map_ss():
    """the reference's sequence-wide rich-map script on a four-frame synthetic sequence."""
    import yaml
    from pcl_augmentation_b200 import synth_io
    from pcl_augmentation_b200.semantic_segmentation.Real3DAug.tools.datasets import SemanticKITTI
    from tests.helpers import RICH_MAP_SS_SEEDS, rich_map_ss_cases
    t0 = time.time()
    cases = rich_map_ss_cases()
    root = tempfile.mkdtemp(prefix="r3d_golden_")
    try:
        cwd, _, cfg = synth_io.write_ss_dataset(cases, root)
        cfg["path"]["maps_path"] = os.path.join(root, "maps", "small", "npz")      # the script pops three components
        with open(os.path.join(root, "config/semantic-kitti.yaml"), "w") as f:
            yaml.safe_dump(cfg, f)
        shim.run_rich_map_ss(cwd, ["1", "00", "no"])
        z = np.load(os.path.join(root, "maps/small/npz/00.npz"))
        m, move = z["map"], z["move"]
        assert set(np.unique(m)) <= {0.0, 1.0, 2.0, 3.0}
        poses = np.loadtxt(os.path.join(root, "data/sequences/00/poses.txt"))
        ds = SemanticKITTI.__new__(SemanticKITTI)
        rec = {"meta": json.dumps(dict(seeds=list(RICH_MAP_SS_SEEDS))), "map_dtype": str(m.dtype), "map": m.astype(np.uint8),
               "move": move, "poses": np.stack([ds.create_transform_matrix(poses, i) for i in range(len(cases))]),
               "placement_labels": json.dumps({int(k): [int(x) for x in v] for k, v in cfg["insertion"]["placement_labels"].items()})}
        for i, c in enumerate(cases):
            rec[f"digest{i}"] = synth.case_digest(c)
        np.savez_compressed(os.path.join(GOLDEN_DIR, "rich_map_ss.npz"), **rec)
        print(f"rich_map_ss: {time.time() - t0:.1f}s map", m.shape, m.dtype, "move", move.ravel(),
              "cells per class", [int((m == v).sum()) for v in (1, 2, 3)])
    finally:
        shutil.rmtree(root, ignore_errors=True)


def _pack_samples(rec, prefix, samples):
    """{name: (anno, pcl)} -> flat arrays (names joined, annotation strings, point counts, concatenated float64 rows)."""
    names = sorted(samples)
    rec[prefix + "_names"] = json.dumps(names)
    rec[prefix + "_annos"] = json.dumps([samples[n][0] for n in names])
    rec[prefix + "_counts"] = np.array([len(samples[n][1]) for n in names], dtype=np.int64)
    rec[prefix + "_pcl"] = (np.concatenate([samples[n][1] for n in names]) if names else np.zeros((0, 5)))
    assert rec[prefix + "_pcl"].dtype == np.float64


def gen_cut_objects_od():
    """the reference's object_cut_out.py on three synthetic frames holding 36 annotated objects."""
    import yaml
    from pcl_augmentation_b200 import synth_io
    from tests.helpers import cut_object_cases, write_kitti_camera_files, read_sample_dir
    t0 = time.time()
    cases = cut_object_cases("od")
    root = tempfile.mkdtemp(prefix="r3d_golden_")
    try:
        cwd, _, cfg = synth_io.write_od_dataset(cases, root)
        write_kitti_camera_files(root, len(cases))
        cfg["path"]["sample_path"] = os.path.join(root, "cut")
        os.makedirs(cfg["path"]["sample_path"])
        with open(os.path.join(root, "config/KITTI.yaml"), "w") as f:
            yaml.safe_dump(cfg, f)
        shim.run_cut_objects_od(cwd)
        rec = {"meta": json.dumps(dict(task="od"))}
        total = 0
        for i, c in enumerate(cases):
            rec[f"digest{i}"] = synth.array_digest(c.pcl5)
        for cls in cfg["insertion"]["classes"]:
            samples = read_sample_dir(os.path.join(root, "cut", cls))
            _pack_samples(rec, cls, samples)
            total += len(samples)
        n_lines = sum(len(c.box_lines) for c in cases)
        np.savez_compressed(os.path.join(GOLDEN_DIR, "cut_objects_od.npz"), **rec)
        print(f"cut_objects_od: {time.time() - t0:.1f}s {total} samples saved of {n_lines} annotation lines")
    finally:
        shutil.rmtree(root, ignore_errors=True)


def gen_cut_objects_ss():
    """cut_out.py, then filter_objects.py, on a three-frame synthetic sequence."""
    import yaml
    from pcl_augmentation_b200 import synth_io
    from tests.helpers import cut_object_cases, read_sample_dir
    t0 = time.time()
    cases = cut_object_cases("ss")
    root = tempfile.mkdtemp(prefix="r3d_golden_")
    try:
        cwd, _, cfg = synth_io.write_ss_dataset(cases, root)
        cfg["path"]["bbox_path"] = os.path.join(root, "cut")
        with open(os.path.join(root, "config/semantic-kitti.yaml"), "w") as f:
            yaml.safe_dump(cfg, f)
        shim.run_cut_objects_ss(cwd, ["1", 0, "no"])       # the sequence prompt must yield an int, see shim
        rec = {"meta": json.dumps(dict(task="ss"))}
        for i, c in enumerate(cases):
            rec[f"digest{i}"] = synth.array_digest(c.pcl5)
        total = 0
        folders = [cfg["labels"][c] for c in cfg["insertion"]["classes"]]
        for folder in folders:
            samples = read_sample_dir(os.path.join(root, "cut", folder))
            _pack_samples(rec, folder, samples)
            total += len(samples)
        shim.run_filter_objects_ss(cwd, ["1", 0, "no"])
        kept = 0
        for folder in folders:
            names = sorted(read_sample_dir(os.path.join(root, "cut", folder)))
            rec[folder + "_kept"] = json.dumps(names)
            kept += len(names)
        np.savez_compressed(os.path.join(GOLDEN_DIR, "cut_objects_ss.npz"), **rec)
        print(f"cut_objects_ss: {time.time() - t0:.1f}s {total} samples saved, {kept} left after filter_objects")
    finally:
        shutil.rmtree(root, ignore_errors=True)


GENERATORS = {
    "cut_objects_od": gen_cut_objects_od,
    "cut_objects_ss": gen_cut_objects_ss,
    "rich_map_od": gen_rich_map_od,
    "rich_map_ss": gen_rich_map_ss,
    "fn_projection": gen_fn_projection,
    "fn_projection_full": gen_fn_projection_full,
    "fn_cut_bbox": gen_fn_cut_bbox,
    "fn_places_od": lambda: gen_fn_places("od"),
    "fn_places_ss": lambda: gen_fn_places("ss"),
}
for _n, _s in E2E_CASES.items():
    GENERATORS[_n] = (lambda n=_n, s=_s: gen_e2e(n, s))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*")
    args = ap.parse_args()
    assert shim.available(), "needs the reference at " + shim.REFERENCE_ROOT
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, fn in GENERATORS.items():
        if args.only and name not in args.only:
            continue
        fn()


if __name__ == "__main__":
    main()
