"""CPU tier: dataset adapters (reference file formats in and out), schedule drawing and the frame-marker logic of the
dataset driver.  The engine is replaced by an oracle-backed stand-in (tests/dataset_helpers.py) so that the HOST logic
is checked against what the unmodified reference's insertion.py wrote (tests/golden/e2e_*.npz)."""
import os
import random

import numpy as np
import pytest

from pcl_augmentation_b200 import dataset_driver as drv
from pcl_augmentation_b200 import synth_io
from pcl_augmentation_b200.object_detection.Real3DAug.tools.datasets import KITTI
from pcl_augmentation_b200.semantic_segmentation.Real3DAug.tools.datasets import SemanticKITTI
from tests.dataset_helpers import OracleEngine, PredrawnShuffle, check_outputs_against_golden
from tests.helpers import case_from_golden, load_golden


def test_kitti_adapter_reads_and_writes_reference_formats(tmp_path):
    g = load_golden("e2e_od_a")
    _, case = case_from_golden(g)
    _, out, cfg = synth_io.write_od_dataset([case], str(tmp_path), fixed_counts=case.schedule.counts)
    ds = KITTI(cfg)
    assert len(ds) == 1
    pcl, anno, instance, calib, img = ds[0]                                   # od/ds:56-71
    assert pcl.dtype == np.float64 and pcl.shape == (len(case.pcl5), 5)
    np.testing.assert_array_equal(pcl, np.hstack((case.pcl5[:, :4].astype(np.float32), case.pcl5[:, 4:5])))
    assert anno.endswith("label_2/000000.txt") and instance.shape == (len(pcl), 1)
    folder, number = ds.create_directories("chosen")
    assert (folder, number) == ("chosen/00", 0)
    for sub in ("velodyne", "check", "label_2", "added_objects"):
        assert os.path.isdir(os.path.join(out, sub))
    assert "2x   Cyclist" in open(os.path.join(out, "setting.txt")).read()
    rows9 = np.full((3, 9), -1.0)
    rows9[:, 0:3] = [[1.5, 2.25, -1.0], [0.1, 0.2, 0.3], [7, 8, 9]]
    rows9[:, 6] = [0.5, 0.25, 0.125]
    ds.save_data(rows9, rows9[:1], folder, "000000", 0, ["Pedestrian 0 0 0 0 0 0 0 1 1 1 1 1 1 0\n"])   # od/ds:76-95
    assert len(ds) == 0                                                        # save_data deletes the item
    np.testing.assert_array_equal(np.fromfile(os.path.join(out, "velodyne/000000.bin"), np.float32).reshape(-1, 4),
                                  np.hstack((rows9[:, 0:3], rows9[:, 6:7])).astype(np.float32))
    lines = open(os.path.join(out, "label_2/000000.txt")).read().splitlines()
    assert lines[:len(case.box_lines)] == case.box_lines and lines[-1].startswith("Pedestrian")
    with pytest.raises(ValueError):
        ds.create_directories("chosen", folder_number=100)


def test_semantic_kitti_adapter_pose_and_files(tmp_path):
    g = load_golden("e2e_ss_a")
    _, case = case_from_golden(g)
    _, out, cfg = synth_io.write_ss_dataset([case], str(tmp_path), fixed_counts=case.schedule.counts)
    ds = SemanticKITTI(cfg, "00")
    pcl, pose, anno, instance, seq = ds[0]                                     # ss/ds:45-62
    assert seq == "00" and pcl.shape == (len(case.pcl5), 5)
    np.testing.assert_array_equal(pose, g["used_pose"])                        # bit-identical to the reference's matrix
    folder, _ = ds.create_directories("chosen")
    assert folder == "chosen/00/sequences"
    for s in cfg["split"]["train"]:
        assert os.path.isdir(os.path.join(cfg["path"]["output_path"], folder, f"{s:02d}", "labels"))
    rows9 = np.full((2, 9), -1.0)
    rows9[:, 0:3] = [[1, 2, 3], [4, 5, 6]]; rows9[:, 6] = [0.5, 0.75]; rows9[:, 7] = [30, 40]
    ds.save_data(rows9, rows9[:1], f"{folder}/00", "000000", 0)                # ss/ds:72-91
    np.testing.assert_array_equal(np.fromfile(os.path.join(out, "labels/000000.label"), np.uint32), [30, 40])
    np.testing.assert_array_equal(np.fromfile(os.path.join(out, "check/000000.bin"), np.float32), [1, 2, 3, 0.5, 30])


def test_draw_schedule_uses_the_reference_rng_sources():
    cfg = {"insertion": {"random": True, "classes": ["a", "b", "c"], "number_of_object": 7}}
    random.seed(3); np.random.seed(3)
    counts, perms = drv.draw_schedule(cfg, [120, 5, 100])
    random.seed(3); np.random.seed(3)
    counts2, perms2 = drv.draw_schedule(cfg, [120, 5, 100])
    np.testing.assert_array_equal(counts, counts2); np.testing.assert_array_equal(perms, perms2)
    assert counts.sum() == 7 and perms.shape == (8, 3, 100)
    assert sorted(perms[0, 1][perms[0, 1] >= 0]) == [0, 1, 2, 3, 4] and (perms[0, 1, 5:] == -1).all()
    assert len(set(perms[2, 0])) == 100 and perms[2, 0].max() < 120
    cfg["insertion"].update(random=False, number_of_classes=[1, 0, 2])
    counts, perms = drv.draw_schedule(cfg, [120, 5, 100])
    assert list(counts) == [1, 0, 2] and perms.shape[0] == 4


@pytest.mark.parametrize("name", ["e2e_od_a", "e2e_ss_b"])
def test_driver_host_logic_matches_reference_run(tmp_path, name):
    """Whole driver (adapters, markers, schedule tables, writers) with the oracle in the engine's place: the output
    files equal what the reference's own insertion.py wrote, byte for byte where the reference is deterministic."""
    g = load_golden(name)
    spec, case = case_from_golden(g)
    task = spec["task"]
    write = synth_io.write_od_dataset if task == "od" else synth_io.write_ss_dataset
    _, out, cfg = write([case], str(tmp_path), fixed_counts=case.schedule.counts)
    with PredrawnShuffle(case.schedule.perms, len(cfg["insertion"]["classes"])):
        if task == "od":
            folder, written, skipped = drv.augment_kitti(cfg, batch_size=4, engine_cls=OracleEngine, log=lambda *a: None)
        else:
            folder, written, skipped = drv.augment_semantic_kitti(cfg, "00", batch_size=4, engine_cls=OracleEngine,
                                                                  log=lambda *a: None)
    assert (written, skipped) == (1, 0) and os.path.join(cfg["path"]["output_path"], folder) == out
    check_outputs_against_golden(g, case, out, task)
    # a second run finds the marker and leaves the frame alone (od/ins:335-338)
    before = os.path.getmtime(os.path.join(out, "velodyne/000000.bin"))
    run = drv.augment_kitti if task == "od" else (lambda c, **k: drv.augment_semantic_kitti(c, "00", **k))
    _, written, skipped = run(cfg, batch_size=4, engine_cls=OracleEngine, log=lambda *a: None)
    assert (written, skipped) == (0, 0) and os.path.getmtime(os.path.join(out, "velodyne/000000.bin")) == before


def test_driver_removes_marker_when_nothing_was_inserted(tmp_path):
    g = load_golden("e2e_od_b")
    _, case = case_from_golden(g)
    _, out, cfg = synth_io.write_od_dataset([case], str(tmp_path), fixed_counts=[0, 0])       # nothing requested
    _, written, skipped = drv.augment_kitti(cfg, engine_cls=OracleEngine, log=lambda *a: None)
    assert (written, skipped) == (0, 1)
    assert not os.path.exists(os.path.join(out, "added_objects/000000.txt"))                   # od/ins:616-620
    assert not os.path.exists(os.path.join(out, "velodyne/000000.bin"))


def test_failed_frame_is_given_back_and_the_others_are_written(tmp_path):
    """A scan that ends with a non-zero status (what the reference raises as an exception) must not take the rest of
    its batch with it: the other frames are written, the failed frame's marker is removed (so a later run retries it)
    and the failure is reported after the run."""
    ga, gb = load_golden("e2e_od_a"), load_golden("e2e_od_b")
    (_, ca), (_, cb) = case_from_golden(ga), case_from_golden(gb)
    _, out, cfg = synth_io.write_od_dataset([ca, cb, ca], str(tmp_path), fixed_counts=ca.schedule.counts)

    class FailingSecond(OracleEngine):
        def augment_batch(self, scans):
            res = super().augment_batch(scans)
            res[1].status = -4                      # e.g. IndexError: class list shorter than a window (od/ins:410)
            return res

    with PredrawnShuffle([ca.schedule.perms, ca.schedule.perms, ca.schedule.perms], len(cfg["insertion"]["classes"])):
        with pytest.raises(drv.FrameError) as err:
            drv.augment_kitti(cfg, batch_size=8, engine_cls=FailingSecond, log=lambda *a: None)
    assert err.value.failures == [("000001", -4)]
    assert os.path.exists(os.path.join(out, "velodyne/000000.bin")) and os.path.exists(os.path.join(out, "velodyne/000002.bin"))
    assert not os.path.exists(os.path.join(out, "added_objects/000001.txt"))
    assert not os.path.exists(os.path.join(out, "velodyne/000001.bin"))


def test_waymo_adapter_and_config(tmp_path):
    """The `Waymo` adapter on frames in the reference's converted layout (ss/ds:218-410) and the shipped waymo.yaml:
    LiDAR mounting shift in / out, pose correction, per-sequence frame lists, ground labels for addjust_map_2."""
    import yaml
    from pcl_augmentation_b200 import engine as eng_mod
    from pcl_augmentation_b200.semantic_segmentation.Real3DAug.tools.datasets import Waymo
    cfg_path = os.path.join(os.path.dirname(eng_mod.__file__), "config", "waymo.yaml")
    with open(cfg_path) as f:
        cfg = yaml.safe_load(f)
    assert cfg["insertion"]["classes"] == [5, 6, 7] and cfg["insertion"]["placement_labels"] == {1: [18, 19], 2: [22], 3: [20]}
    assert cfg["insertion"]["road_indexes"] == [18, 19, 20, 22] and cfg["labels"][22] == "sidewalk"
    rng = np.random.default_rng(5)
    data, out = tmp_path / "data", tmp_path / "out"
    frames = {}
    for seq, names in (("seg_a", ["0000", "0001"]), ("seg_b", ["0000"])):
        for sub in ("lidar", "labels_v3_2", "poses"):
            os.makedirs(data / seq / sub)
        for n in names:
            pts = rng.uniform(-20, 20, (50, 6)).astype(np.float32)
            lab = np.stack((rng.integers(0, 9, 50), rng.integers(0, 23, 50)), axis=1).astype(np.int32)
            pose = np.eye(4); pose[:3, 3] = rng.uniform(-5, 5, 3)
            np.save(data / seq / "lidar" / f"{n}.npy", pts); np.save(data / seq / "labels_v3_2" / f"{n}.npy", lab)
            np.save(data / seq / "poses" / f"{n}.npy", pose)
            frames[(seq, n)] = (pts, lab, pose)
    cfg["path"].update(dataset_path=str(data), annotation_path=str(tmp_path / "anno"), output_path=str(out))
    ds = Waymo(cfg)
    assert ds.sequence_names == ["seg_a", "seg_b"] and len(ds) == 3 and ds.sequence == "seg_a"
    xyzi, labels, pose, anno, name = ds.read_frame(2)
    pts, lab, p0 = frames[("seg_b", "0000")]
    np.testing.assert_array_equal(xyzi[:, :3], (pts[:, :3].astype(np.float64) - Waymo.LiDAR_location).astype(np.float32))   # ss/ds:262-267
    np.testing.assert_array_equal(labels, lab[:, 1].astype(np.uint32))
    corr = np.eye(4); corr[:3, 3] = Waymo.LiDAR_location
    np.testing.assert_array_equal(pose, p0 @ corr)
    assert anno == f"{tmp_path}/anno/seg_b/bbox/0000.txt" and name == "0000"
    folder, number = ds.create_directories("chosen")
    assert (folder, number) == ("chosen/00", 0) and os.path.isdir(out / folder / "seg_a" / "labels_v3_2")
    ds._write(f"{folder}/seg_a", "0000", xyzi, labels, np.hstack((xyzi[:2], [[5.0], [6.0]])).astype(np.float32))
    back = np.load(out / folder / "seg_a" / "lidar" / "0000.npy")
    np.testing.assert_allclose(back[:, :3], pts[:, :3], rtol=0, atol=4e-6)                     # shifted back (ss/ds:286-287)
    assert np.load(out / folder / "seg_a" / "labels_v3_2" / "0000.npy").shape == (50, 1)
