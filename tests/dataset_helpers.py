"""Test infrastructure for the dataset-driver tests: pre-drawn shuffles and an ORACLE-backed stand-in engine
(CPU tier only; the GPU tier runs the real CUDA engine through the same driver)."""
import hashlib
import os
import random

import numpy as np

from oracle import real3d_oracle as orc
from pcl_augmentation_b200.engine import ScanResult


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class PredrawnShuffle:
    """Stands in for ``dataset_driver._permutation_heads``: frame f gets the pre-drawn table of the golden run
    ([event][class][try]; a head shorter than a window is followed by the remaining indices in order, as a shuffled list
    would be).  ``perms``: one table, or a list with one table per frame."""

    def __init__(self, perms, n_classes):
        self.tables = list(perms) if isinstance(perms, (list, tuple)) else [perms]
        self.n_classes, self.frame = n_classes, 0

    def __call__(self, n_events, list_lens, tries):
        table = np.asarray(self.tables[self.frame % len(self.tables)])
        self.frame += 1
        assert table.shape[0] == n_events and table.shape[1] == len(list_lens) == self.n_classes
        out = np.full((n_events, len(list_lens), tries), -1, dtype=np.int32)
        for e in range(n_events):
            for c, n in enumerate(list_lens):
                head = [int(v) for v in table[e][c] if v >= 0]
                hs = set(head)
                order = (head + [j for j in range(n) if j not in hs])[:min(tries, n)]
                out[e, c, :len(order)] = order
        return out

    def __enter__(self):
        from pcl_augmentation_b200 import dataset_driver as drv
        self.old = drv._permutation_heads
        drv._permutation_heads = self
        return self

    def __exit__(self, *exc):
        from pcl_augmentation_b200 import dataset_driver as drv
        drv._permutation_heads = self.old


class OracleEngine:
    """Real3DEngine's constructor / augment_batch / close surface on top of the numpy oracle (CPU tests of the host
    logic only — never part of the product path)."""

    def __init__(self, task, config, db, *, max_scans, max_points, yaw_steps=360, map_data=None, **kw):
        self.task, self.config, self.db, self.yaw_steps, self.map_data = task, config, db, yaw_steps, map_data
        self.max_scans, self.max_points = max_scans, max_points

    def augment_batch(self, scans):
        assert len(scans) <= self.max_scans
        out = []
        for s in scans:
            assert len(s.xyzi) <= self.max_points
            pcl5 = np.hstack((np.asarray(s.xyzi, dtype=np.float64), np.asarray(s.labels, dtype=np.float64).reshape(-1, 1)))
            ref = orc.augment_scan(self.task, pcl5, [l.rstrip("\n") for l in s.box_lines], self.db, s.counts, s.perms,
                                   self.config, maps=s.maps, map_data=self.map_data, transform_matrix=s.pose,
                                   mode="closed", yaw_steps=self.yaw_steps)
            arr = orc.save_arrays(self.task, ref)
            out.append(ScanResult(velodyne=arr["velodyne"], labels=arr["labels"].ravel() if self.task == "ss" else
                                  np.zeros(len(arr["velodyne"]), np.uint32), check=arr["check"],
                                  inserted=[(n, int(r), c) for n, r, c in ref["inserted"]], lines=list(ref["lines"]),
                                  boxes=[], visible=[]))
        return out

    def close(self):
        pass


def check_outputs_against_golden(g, case, out_dir, task, frame="000000", exact_tail=False):
    """The files the driver wrote vs what the unmodified reference's insertion.py wrote for the same inputs."""
    with open(os.path.join(out_dir, "added_objects", f"{frame}.txt")) as f:
        assert f.read() == str(g["inserted"])                                                  # placement choices
    velodyne = np.fromfile(os.path.join(out_dir, "velodyne", f"{frame}.bin"), dtype=np.float32).reshape(-1, 4)
    check = np.fromfile(os.path.join(out_dir, "check", f"{frame}.bin"), dtype=np.float32).reshape(-1, 5 if task == "ss" else 4)
    n0 = len(case.pcl5)
    keep = np.unpackbits(g["keep_orig"])[:n0].astype(bool)
    n_kept = int(g["n_kept"])
    assert len(velodyne) == int(g["n_out"])                                                    # point counts
    np.testing.assert_array_equal(velodyne[:n_kept], case.pcl5[keep][:, :4].astype(np.float32))   # keep-mask, bit exact
    np.testing.assert_allclose(velodyne[n_kept:], g["tail"], rtol=0, atol=1e-6)                # xyz within 1e-6 m
    np.testing.assert_allclose(check, g["check"], rtol=0, atol=1e-6)
    if exact_tail:
        assert sha(velodyne) == str(g["velodyne_sha"])
    if task == "ss":
        labels = np.fromfile(os.path.join(out_dir, "labels", f"{frame}.label"), dtype=np.uint32)
        assert sha(labels) == str(g["labels_sha"])
    else:
        with open(os.path.join(out_dir, "label_2", f"{frame}.txt")) as f:
            assert f.read() == str(g["label_2"])
    return velodyne
