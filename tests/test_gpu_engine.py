"""GPU tier: the batched CUDA engine (through the C ABI) against (a) what the unmodified reference's insertion.py
wrote for the golden cases and (b) the numpy oracle on fresh seeded scans."""
import hashlib

import numpy as np
import pytest

from oracle import real3d_oracle as orc
from pcl_augmentation_b200 import synth
from pcl_augmentation_b200.engine import Real3DEngine, scan_input_from_case
from tests.helpers import GOLDEN_SHAPE, case_from_golden, load_golden, parse_inserted

pytestmark = pytest.mark.gpu

E2E = ["e2e_od_a", "e2e_od_b", "e2e_ss_a", "e2e_ss_b"]


def make_engine(case, n_scans=1, max_points=None, **kw):
    return Real3DEngine(case.task, case.config, case.db, max_scans=n_scans,
                        max_points=max_points or len(case.pcl5), map_data=case.map_data, **kw)


def oracle_run(case, pose=None, **kw):
    res = orc.augment_scan(case.task, case.pcl5, case.box_lines, case.db, case.schedule.counts, case.schedule.perms,
                           case.config, maps=case.maps, map_data=case.map_data,
                           transform_matrix=case.pose if pose is None else pose, mode="closed", **kw)
    return res, orc.save_arrays(case.task, res)


def assert_matches_oracle(case, got, ref, want, exact_tail=False):
    assert [(n, int(r)) for n, r, _ in got.inserted] == [(n, int(r)) for n, r, _ in ref["inserted"]]
    assert got.velodyne.shape == want["velodyne"].shape
    n_keep = int(ref["keep_orig"].sum())
    np.testing.assert_array_equal(got.velodyne[:n_keep], want["velodyne"][:n_keep])       # keep-mask, bit exact
    np.testing.assert_allclose(got.velodyne[n_keep:], want["velodyne"][n_keep:], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got.check, want["check"], rtol=0, atol=1e-6)
    if exact_tail:
        np.testing.assert_array_equal(got.velodyne, want["velodyne"])
    if case.task == "ss":
        np.testing.assert_array_equal(got.labels, want["labels"].ravel())
    else:
        assert got.lines == ref["lines"]


@pytest.mark.parametrize("name", E2E)
def test_engine_matches_reference_run(name):
    """Same seeded inputs and pre-drawn shuffles as the reference's own insertion.py run."""
    g = load_golden(name)
    spec, case = case_from_golden(g)
    pose = g["used_pose"] if "used_pose" in g.files else None
    eng = make_engine(case)
    got = eng.augment_batch([scan_input_from_case(case, pose)])[0]
    eng.close()
    assert [(n, r) for n, r, _ in got.inserted] == parse_inserted(str(g["inserted"]))          # placement choices
    n0 = len(case.pcl5)
    keep = np.unpackbits(g["keep_orig"])[:n0].astype(bool)
    n_kept = int(g["n_kept"])
    assert len(got.velodyne) == int(g["n_out"])                                                # point counts
    np.testing.assert_array_equal(got.velodyne[:n_kept], case.pcl5[keep][:, :4].astype(np.float32))   # keep-mask
    np.testing.assert_allclose(got.velodyne[n_kept:], g["tail"], rtol=0, atol=1e-6)            # xyz within 1e-6 m
    np.testing.assert_allclose(got.check, g["check"], rtol=0, atol=1e-6)
    if case.task == "ss":
        np.testing.assert_array_equal(got.labels[n_kept:], g["tail_labels"])
        assert hashlib.sha256(np.ascontiguousarray(got.labels).tobytes()).hexdigest() == str(g["labels_sha"])
    else:
        ref_lines = str(g["label_2"]).splitlines(keepends=True)
        assert ref_lines[len(case.box_lines):] == got.lines


@pytest.mark.parametrize("task,seed,counts", [("od", 101, [2, 1]), ("od", 102, [0, 3]), ("ss", 201, [1, 1, 0, 1, 0, 1]),
                                              ("ss", 202, [0, 0, 1, 2, 0, 0])])
def test_engine_vs_oracle_fresh_seeds(task, seed, counts):
    case = synth.make_case(task, seed, shape=GOLDEN_SHAPE, counts=counts, obj_range=(4.0, 16.0))
    eng = make_engine(case)
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    eng.close()
    ref, want = oracle_run(case)
    assert_matches_oracle(case, got, ref, want)


def test_engine_ragged_batch_od():
    """Several scans of different sizes in one batch, processed in lock-step rounds."""
    shapes = [synth.ScanShape(32, 600, 2.0, -24.8), synth.ScanShape(24, 500, 2.0, -24.8), synth.ScanShape(32, 450, 2.0, -24.8)]
    cases = [synth.make_case("od", 300 + i, shape=sh, counts=c, obj_range=(4.0, 16.0), db_seed=0)
             for i, (sh, c) in enumerate(zip(shapes, ([1, 2], [2, 0], [0, 1])))]
    # one shared cut-object database (replicated per GPU in production): use the first case's db for all
    for c in cases[1:]:
        c.db = cases[0].db
    eng = Real3DEngine("od", cases[0].config, cases[0].db, max_scans=4, max_points=max(len(c.pcl5) for c in cases))
    got = eng.augment_batch([scan_input_from_case(c) for c in cases])
    for c, g in zip(cases, got):
        ref, want = oracle_run(c)
        assert_matches_oracle(c, g, ref, want)
    # re-arm the resident batch and run again: identical results
    eng.reset()
    eng.run()
    again = eng.unpack(eng.fetch_raw())
    for a, g in zip(again, got):
        np.testing.assert_array_equal(a.velodyne, g.velodyne)
        assert a.inserted == g.inserted
    eng.close()


def test_engine_yaw_steps_1024_and_image_64x2048():
    case = synth.make_case("od", 401, shape=GOLDEN_SHAPE, counts=[1, 1], obj_range=(4.0, 16.0))
    eng = make_engine(case, rows=64, cols=2048, yaw_steps=1024)
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    eng.close()
    ref, want = oracle_run(case, num_row=64, num_column=2048, yaw_steps=1024)
    assert_matches_oracle(case, got, ref, want)


def test_engine_nothing_to_insert_and_impossible_objects():
    case = synth.make_case("od", 501, shape=GOLDEN_SHAPE, counts=[0, 0], obj_range=(4.0, 16.0))
    eng = make_engine(case)
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    assert got.inserted == [] and len(got.velodyne) == len(case.pcl5)
    np.testing.assert_array_equal(got.velodyne, case.pcl5[:, :4].astype(np.float32))
    eng.close()
    # objects far away (few points): windows exhaust, remaining counts run down like the reference's time-out logic
    far = synth.make_case("od", 502, shape=GOLDEN_SHAPE, counts=[1, 1], obj_range=(30.0, 40.0), db_seed=3, tries=10)
    eng = make_engine(far, max_tries=10)
    got = eng.augment_batch([scan_input_from_case(far)])[0]
    eng.close()
    ref, want = oracle_run(far, max_tries=10)
    assert_matches_oracle(far, got, ref, want)


def test_candidate_flags_match_oracle_first_try():
    case = synth.make_case("od", 601, shape=GOLDEN_SHAPE, counts=[1, 0], obj_range=(4.0, 16.0))
    trace = []
    ref, want = oracle_run(case, trace=trace)
    tries = [t for t in trace if t[0] == "try"]
    eng = make_engine(case)
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    flags, level, vis = eng.debug_candidates(0)
    eng.close()
    assert_matches_oracle(case, got, ref, want)
    if len(ref["inserted"]) == 1 and len(tries) >= 1:
        last = tries[-1]
        feasible = [k for k in range(1, 361) if flags[k] == 3]
        assert feasible == last[2]


@pytest.mark.parametrize("task,seed,counts", [("od", 701, [2, 2]), ("ss", 702, [1, 1, 1, 1, 0, 0])])
def test_in_place_image_patch_equals_full_reprojection(task, seed, counts):
    """The incremental slot update (window mask + z-buffer patch + windowed close/fill) must give exactly what a full
    re-projection of every slot gives."""
    case = synth.make_case(task, seed, shape=GOLDEN_SHAPE, counts=counts, obj_range=(4.0, 16.0))
    outs = []
    for force in (False, True):
        eng = make_engine(case, force_full_projection=force)
        outs.append(eng.augment_batch([scan_input_from_case(case)])[0])
        st = eng.stats()
        assert (st["patched_scans"] == 0) == force
        eng.close()
    a, b = outs
    assert a.inserted == b.inserted and len(a.inserted) >= 2
    np.testing.assert_array_equal(a.velodyne, b.velodyne)
    np.testing.assert_array_equal(a.labels, b.labels)
    np.testing.assert_array_equal(a.check, b.check)
