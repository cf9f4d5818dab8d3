"""GPU tier: the BASELINE.json configurations at their FULL sizes.

C1  one KITTI-shape HDL-64 scan (120 000 pts), 10 cut pedestrians / cyclists            -> vs the oracle, everything
C2  SemanticKITTI-shape scan (124 992 pts), 64 x 2048 range image, 20 rare-class objects -> vs the oracle, everything
C3  batch of 256 KITTI-shape scans, 1024 yaw candidates, sub-batch streams               -> oracle on a sample + batch
    independence (a scan's result does not depend on its slot, its neighbours or the number of sub-batches)
C4  OS1-128-shape scan (262 144 pts), 128 x 2048 image, 50 objects                       -> vs the oracle, everything
C5  a stream of batches through ScanPipeline (H2D / compute / D2H overlapped)            -> equals batch-by-batch runs
Bar: keep-masks, placement choices, point counts, labels, annotation lines exact; inserted xyz within 1e-6 m.
"""
import numpy as np
import pytest

from oracle import real3d_oracle as orc
from pcl_augmentation_b200 import synth
from pcl_augmentation_b200.engine import Real3DEngine, scan_input_from_case
from pcl_augmentation_b200.pipeline import ScanPipeline

pytestmark = pytest.mark.gpu


def oracle(case, **kw):
    ref = orc.augment_scan(case.task, case.pcl5, case.box_lines, case.db, case.schedule.counts, case.schedule.perms,
                           case.config, maps=case.maps, map_data=case.map_data, transform_matrix=case.pose,
                           mode="closed", **kw)
    return ref, orc.save_arrays(case.task, ref)


def assert_same(case, got, ref, want):
    assert got.status == 0
    assert [(n, int(r)) for n, r, _ in got.inserted] == [(n, int(r)) for n, r, _ in ref["inserted"]]     # choices
    assert got.velodyne.shape == want["velodyne"].shape                                                   # counts
    n_keep = int(ref["keep_orig"].sum())
    np.testing.assert_array_equal(got.velodyne[:n_keep], want["velodyne"][:n_keep])                       # keep-mask
    np.testing.assert_allclose(got.velodyne[n_keep:], want["velodyne"][n_keep:], rtol=0, atol=1e-6)       # 1e-6 m
    np.testing.assert_allclose(got.check, want["check"], rtol=0, atol=1e-6)
    if case.task == "ss":
        np.testing.assert_array_equal(got.labels, want["labels"].ravel())
    else:
        assert got.lines == ref["lines"]


def same_result(a, b):
    return (a.inserted == b.inserted and np.array_equal(a.velodyne, b.velodyne) and np.array_equal(a.check, b.check)
            and np.array_equal(a.labels, b.labels) and a.lines == b.lines)


def test_c1_kitti_scan_10_objects():
    case = synth.make_case("od", 4001, number_of_object=10)
    assert len(case.pcl5) == 120000
    eng = Real3DEngine("od", case.config, case.db, max_scans=1, max_points=len(case.pcl5))
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    eng.close()
    ref, want = oracle(case)
    assert len(ref["inserted"]) >= 8
    assert_same(case, got, ref, want)


def test_c2_semantickitti_scan_64x2048_20_objects():
    case = synth.make_case("ss", 4002, shape=synth.SEMKITTI_SHAPE, number_of_object=20)
    assert len(case.pcl5) == 124992
    eng = Real3DEngine("ss", case.config, case.db, max_scans=1, max_points=len(case.pcl5), rows=64, cols=2048,
                       map_data=case.map_data)
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    eng.close()
    ref, want = oracle(case, num_row=64, num_column=2048)
    assert len(ref["inserted"]) >= 10
    assert_same(case, got, ref, want)


def test_c4_os128_scan_128x2048_50_objects():
    case = synth.make_case("od", 4004, shape=synth.OS128_SHAPE, number_of_object=50)
    assert len(case.pcl5) == 262144
    eng = Real3DEngine("od", case.config, case.db, max_scans=1, max_points=len(case.pcl5), rows=128, cols=2048)
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    eng.close()
    ref, want = oracle(case, num_row=128, num_column=2048)
    assert len(ref["inserted"]) >= 25
    assert_same(case, got, ref, want)


def test_c3_batch_of_256_scans_1024_candidates():
    base = [synth.make_case("od", 4100 + i, number_of_object=10) for i in range(8)]
    cases = []
    for j in range(256):
        c = base[j % 8]
        sched = synth.make_schedule(4200 + j // 8 % 4, 2, 10, [len(c.db[k]) for k in c.config["insertion"]["classes"]])
        cases.append(synth.Case(c.task, c.config, c.pcl5, c.box_lines, c.db, sched, maps=c.maps, cars=c.cars))
    inputs = [scan_input_from_case(c) for c in cases]
    eng = Real3DEngine("od", cases[0].config, cases[0].db, max_scans=256, max_points=120000, yaw_steps=1024, max_events=11)
    res = eng.augment_batch(inputs)
    assert all(r.status == 0 for r in res)
    # (a) oracle on a sample, K = 1024
    for j in (0, 77, 255):
        ref, want = oracle(cases[j], yaw_steps=1024)
        assert_same(cases[j], res[j], ref, want)
    # (b) slots 0..31 and 32..63 hold the same (scan, schedule) pairs (period 32): identical results whatever the
    #     slot / sub-batch a scan lands in
    for j in range(32):
        assert same_result(res[j], res[j + 32]) and same_result(res[j], res[j + 224])
    # (c) independent of the number of concurrent sub-batches and of the neighbours in the batch
    eng.set_sub_batches(1)
    res1 = eng.augment_batch(inputs[:64][::-1])
    eng.close()
    for j in range(64):
        assert same_result(res1[j], res[63 - j])
    # (d) size-independent properties: the output is a subsequence of the input followed by a subsequence of the
    #     `check` record (later insertions may occlude points of earlier ones); counts add up
    def subsequence_len(rows, pool):
        j = 0
        for row in rows:
            while j < len(pool) and pool[j] != row:
                j += 1
            if j == len(pool):
                return None
            j += 1
        return len(rows)
    for j in (5, 130):
        r, c = res[j], cases[j]
        src = [row.tobytes() for row in c.pcl5[:, :4].astype(np.float32)]
        out = [row.tobytes() for row in r.velodyne]
        src_set = set(src)
        n_kept = next((i for i, row in enumerate(out) if row not in src_set), len(out))
        assert subsequence_len(out[:n_kept], src) == n_kept
        chk = [row.tobytes() for row in np.ascontiguousarray(r.check[:, :4])]
        assert subsequence_len(out[n_kept:], chk) == len(out) - n_kept
        assert len(r.check) == sum(r.visible) and len(r.inserted) == len(r.visible) <= 10


def test_c5_stream_through_the_pipeline_equals_batch_runs():
    cases = [synth.make_case("od", 4300 + i, shape=synth.ScanShape(32, 600, 2.0, -24.8), number_of_object=3,
                             obj_range=(4.0, 16.0)) for i in range(12)]
    batches = [[scan_input_from_case(c) for c in cases[i:i + 4]] for i in range(0, 12, 4)] * 2      # 6 batches of 4
    kw = dict(max_scans=4, max_points=len(cases[0].pcl5))
    pipe = ScanPipeline("od", cases[0].config, cases[0].db, depth=3, **kw)
    streamed = pipe.augment_stream(batches)
    pipe.close()
    eng = Real3DEngine("od", cases[0].config, cases[0].db, **kw)
    for got, batch in zip(streamed, batches):
        want = eng.augment_batch(batch)
        assert len(got) == len(want) and all(same_result(a, b) for a, b in zip(got, want))
    eng.close()
    ref, want = oracle(cases[5])
    assert_same(cases[5], streamed[1][1], ref, want)
