"""GPU tier: the dataset drivers (the reference's two insertion.py scripts re-hosted on the CUDA engine) on on-disk
datasets in the reference's formats, against what the unmodified reference wrote for the same inputs."""
import os
import threading

import numpy as np
import pytest

from pcl_augmentation_b200 import dataset_driver as drv
from pcl_augmentation_b200 import synth_io
from tests.dataset_helpers import PredrawnShuffle, check_outputs_against_golden
from tests.helpers import case_from_golden, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["e2e_od_a", "e2e_od_b", "e2e_ss_a", "e2e_ss_b"])
def test_insertion_script_matches_reference_files(tmp_path, name):
    g = load_golden(name)
    spec, case = case_from_golden(g)
    task = spec["task"]
    write = synth_io.write_od_dataset if task == "od" else synth_io.write_ss_dataset
    _, out, cfg = write([case], str(tmp_path), fixed_counts=case.schedule.counts)
    with PredrawnShuffle(case.schedule.perms, len(cfg["insertion"]["classes"])):
        if task == "od":
            folder, written, skipped = drv.augment_kitti(cfg, batch_size=8, log=lambda *a: None)
        else:
            folder, written, skipped = drv.augment_semantic_kitti(cfg, "00", batch_size=8, log=lambda *a: None)
    assert os.path.join(cfg["path"]["output_path"], folder) == out
    if str(g["inserted"]) == "":
        assert (written, skipped) == (0, 1) and not os.path.exists(os.path.join(out, "added_objects/000000.txt"))
        return
    assert (written, skipped) == (1, 0)
    check_outputs_against_golden(g, case, out, task)


def test_multi_frame_batch_with_a_claimed_frame(tmp_path, monkeypatch):
    """Three frames in one engine batch (two different golden scans + a repeat), one of them already claimed by
    "another run": per-frame schedules, per-frame maps, marker skip (od/ins:335-338)."""
    ga, gb = load_golden("e2e_od_a"), load_golden("e2e_od_b")
    (_, ca), (_, cb) = case_from_golden(ga), case_from_golden(gb)
    _, out, cfg = synth_io.write_od_dataset([ca, cb, ca], str(tmp_path), fixed_counts=ca.schedule.counts)
    os.makedirs(os.path.join(out, "added_objects"))
    open(os.path.join(out, "added_objects/000002.txt"), "w").close()           # frame 2 is "in progress" elsewhere
    counts = iter([ca.schedule.counts, cb.schedule.counts])
    monkeypatch.setattr(drv, "generate_seed", lambda config: np.array(next(counts), dtype=np.float64))
    with PredrawnShuffle([ca.schedule.perms, cb.schedule.perms], len(cfg["insertion"]["classes"])):
        _, written, skipped = drv.augment_kitti(cfg, batch_size=8, log=lambda *a: None)
    assert (written, skipped) == (2, 0)
    check_outputs_against_golden(ga, ca, out, "od", frame="000000")
    check_outputs_against_golden(gb, cb, out, "od", frame="000001")
    assert not os.path.exists(os.path.join(out, "velodyne/000002.bin"))
    assert os.path.getsize(os.path.join(out, "added_objects/000002.txt")) == 0


def test_engine_calls_from_another_thread_and_device(tmp_path):
    """The C ABI binds the calling thread to the engine's device (worker threads of ScanPipeline start on device 0):
    create the engine on the LAST visible GPU and drive it from a fresh thread."""
    import torch
    from pcl_augmentation_b200.engine import Real3DEngine, scan_input_from_case
    g = load_golden("e2e_od_a")
    _, case = case_from_golden(g)
    dev = torch.cuda.device_count() - 1
    torch.cuda.set_device(dev)
    try:
        eng = Real3DEngine("od", case.config, case.db, max_scans=1, max_points=len(case.pcl5))
        box = {}
        t = threading.Thread(target=lambda: box.update(res=eng.augment_batch([scan_input_from_case(case)])[0]))
        t.start(); t.join()
        eng.close()
    finally:
        torch.cuda.set_device(0)
    from tests.helpers import parse_inserted
    assert [(n, r) for n, r, _ in box["res"].inserted] == parse_inserted(str(g["inserted"]))
