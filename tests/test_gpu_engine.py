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

E2E = ["e2e_od_a", "e2e_od_b", "e2e_od_c", "e2e_ss_a", "e2e_ss_b", "e2e_ss_c",
       # BASELINE-size scans through the unmodified reference (120 000 / 124 992 points, 112 x 1440) and a class of very
       # large, very close cut objects (7 880 - 15 709 points: global-scratch branches of the candidate selection)
       "e2e_od_full", "e2e_ss_full", "e2e_ss_big"]
# execution models of the engine: the per-scan persistent walker (default) and the staged round kernels
MODES = [pytest.param(False, id="walker"), pytest.param(True, id="staged")]


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


@pytest.mark.parametrize("staged", MODES)
@pytest.mark.parametrize("name", E2E)
def test_engine_matches_reference_run(name, staged):
    """Same seeded inputs and pre-drawn shuffles as the reference's own insertion.py run."""
    g = load_golden(name)
    spec, case = case_from_golden(g)
    pose = g["used_pose"] if "used_pose" in g.files else None
    eng = make_engine(case, staged_rounds=staged)
    got = eng.augment_batch([scan_input_from_case(case, pose)])[0]
    if name == "e2e_ss_big":                       # > 4096 points and a pixel rectangle wider than the shared-memory tile
        assert eng.stats()["select_global"] > 0
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
@pytest.mark.parametrize("staged", MODES)
def test_engine_vs_oracle_fresh_seeds(task, seed, counts, staged):
    case = synth.make_case(task, seed, shape=GOLDEN_SHAPE, counts=counts, obj_range=(4.0, 16.0))
    eng = make_engine(case, staged_rounds=staged)
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
    eng = make_engine(case, candidate_window=False)             # every candidate of the try evaluated, for inspection
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    flags, level, vis = eng.debug_candidates(0)
    eng.close()
    assert_matches_oracle(case, got, ref, want)
    if len(ref["inserted"]) == 1 and len(tries) >= 1:
        last = tries[-1]
        feasible = [k for k in range(1, 361) if flags[k] == 3]
        assert feasible == last[2]


@pytest.mark.parametrize("staged", MODES)
@pytest.mark.parametrize("task,seed,counts", [("od", 701, [2, 2]), ("ss", 702, [1, 1, 1, 1, 0, 0])])
def test_in_place_image_patch_equals_full_reprojection(task, seed, counts, staged):
    """The incremental slot update (window mask + z-buffer patch + windowed close/fill) must give exactly what a full
    re-projection of every slot gives."""
    case = synth.make_case(task, seed, shape=GOLDEN_SHAPE, counts=counts, obj_range=(4.0, 16.0))
    outs = []
    for force in (False, True):
        eng = make_engine(case, force_full_projection=force, staged_rounds=staged)
        outs.append(eng.augment_batch([scan_input_from_case(case)])[0])
        st = eng.stats()
        assert (st["patched_scans"] == 0) == force
        if force and not staged:
            assert st["walker_full_reprojections"] >= len(outs[-1].inserted) - 1 > 0     # every slot after the first one
        eng.close()
    a, b = outs
    assert a.inserted == b.inserted and len(a.inserted) >= 2
    np.testing.assert_array_equal(a.velodyne, b.velodyne)
    np.testing.assert_array_equal(a.labels, b.labels)
    np.testing.assert_array_equal(a.check, b.check)


def test_execution_modes_give_identical_results():
    """Round graphs vs kernel-by-kernel launches, 1 / 5 / 16 concurrent sub-batches and a re-armed second run of the same
    resident batch: the schedule of the kernels must not change a single output byte."""
    base = [synth.make_case("od", 820 + i, shape=GOLDEN_SHAPE, counts=[2, 1], obj_range=(4.0, 16.0)) for i in range(6)]
    cases = [base[i % 6] for i in range(37)]                      # 37 scans: uneven sub-batches
    inputs = [scan_input_from_case(c) for c in cases]
    n_pts = max(len(c.pcl5) for c in cases)
    ref = None
    for graphs, subs in ((None, 0), (False, 1), (True, 1), (True, 5), (True, 16), (False, 16)):      # None: the walker
        eng = Real3DEngine("od", cases[0].config, cases[0].db, max_scans=len(cases), max_points=n_pts,
                           sub_batches=subs, round_graphs=bool(graphs), staged_rounds=graphs is not None)
        staged = eng.stage(inputs)
        eng.load(staged)
        runs = []
        for _ in range(2):                                        # second pass: re-armed resident batch, cached graphs
            eng.run()
            runs.append(eng.unpack(eng.fetch_raw()))
            eng.reset(from_raw_points=len(runs) == 1 and graphs is None)      # the walker also re-arms from the raw points
        eng.close()
        for res in runs:
            assert all(r.status == 0 for r in res) and sum(len(r.inserted) for r in res) >= 37
            if ref is None:
                ref = res
                continue
            for a, b in zip(ref, res):
                assert a.inserted == b.inserted and a.lines == b.lines
                np.testing.assert_array_equal(a.velodyne, b.velodyne)
                np.testing.assert_array_equal(a.check, b.check)
    want_ref, want = oracle_run(cases[3])
    assert_matches_oracle(cases[3], ref[3], want_ref, want)


def test_outputs_on_device_alias_the_fetched_results():
    """The zero-copy device view (for a consumer on the same GPU) holds exactly what fetch copies to the host."""
    import torch
    cases = [synth.make_case("ss", 840 + i, shape=GOLDEN_SHAPE, counts=[1, 1, 0, 0, 0, 0], obj_range=(4.0, 16.0)) for i in range(3)]
    eng = Real3DEngine("ss", cases[0].config, cases[0].db, max_scans=3, max_points=max(len(c.pcl5) for c in cases),
                       map_data=cases[0].map_data)
    eng.load(eng.stage([scan_input_from_case(c) for c in cases]))
    eng.run()
    dev = eng.outputs_on_device()
    assert dev["xyzi"].is_cuda and dev["xyzi"].dtype == torch.float32
    host = eng.unpack(eng.fetch_raw())
    eng_rows = dev["xyzi"].cpu().numpy()
    for s, r in enumerate(host):
        a, b = dev["offsets"][s], dev["offsets"][s + 1]
        np.testing.assert_array_equal(eng_rows[a:b], r.velodyne)
        np.testing.assert_array_equal(dev["labels"][a:b].cpu().numpy().view(np.uint32), r.labels.astype(np.uint32))
        ca, cb = dev["check_offsets"][s], dev["check_offsets"][s + 1]
        np.testing.assert_array_equal(dev["check"][ca:cb].cpu().numpy(), r.check)
    eng.close()



def test_error_conventions_index_capacity_and_arguments():
    """SURVEY §8b error conventions: IndexError where the reference raises it (a class list shorter than a
    MAX_NUM_TRIES window, od/ins:410), Real3DError for capacity / argument problems, per-scan status otherwise."""
    from pcl_augmentation_b200._lib import Real3DError
    # (1) 6 samples per class, nothing placeable (all objects far outside the map): the first window runs past the list
    case = synth.make_case("od", 861, shape=GOLDEN_SHAPE, counts=[1, 0], n_per_class=6, obj_range=(4.0, 16.0))
    for items in case.db.values():
        for _, s in items:
            s["pcl"][:, 0] += 500.0                                  # off the 64 m map: never on the road
    eng = make_engine(case)
    staged = eng.stage([scan_input_from_case(case)])
    eng.load(staged)
    eng.run()
    buf = eng.fetch_raw()
    with pytest.raises(IndexError):
        eng.unpack(buf)
    res = eng.unpack(buf, raise_on_error=False)                       # the scan keeps its points, status says why
    assert res[0].status == -4 and res[0].inserted == [] and len(res[0].velodyne) == len(case.pcl5)
    with pytest.raises(IndexError):                                   # the oracle raises where the reference does
        oracle_run(case)
    eng.close()
    # (2) more scans / points than the engine was built for
    case2 = synth.make_case("od", 862, shape=GOLDEN_SHAPE, counts=[1, 1], obj_range=(4.0, 16.0))
    eng = make_engine(case2, n_scans=1)
    with pytest.raises((Real3DError, AssertionError)):
        eng.stage([scan_input_from_case(case2)] * 2)
    small = Real3DEngine("od", case2.config, case2.db, max_scans=1, max_points=len(case2.pcl5) - 5)
    with pytest.raises(Real3DError):
        small.load(small.stage([scan_input_from_case(case2)]))
    small.close()
    # (3) run before load
    with pytest.raises(Real3DError):
        eng.run()
    eng.close()
    # (4) a schedule that asks for more objects than the engine's event records hold is refused at load time
    tight = Real3DEngine("od", case2.config, case2.db, max_scans=1, max_points=len(case2.pcl5), max_events=2)
    with pytest.raises(Real3DError, match="max_events"):
        tight.load(tight.stage([scan_input_from_case(case2)]))          # counts [1, 1]: 2 objects + 1 > max_events 2
    tight.close()


@pytest.mark.parametrize("staged", MODES)
@pytest.mark.parametrize("task,counts", [("od", [2, 2]), ("ss", [1, 1, 1, 1, 0, 0])])
def test_randomised_sweep_crowded_scenes_vs_oracle(task, counts, staged):
    """16 seeded scans per pipeline with 8 - 15 scene boxes each (the collision pruning — bounding circles, the
    separating-axis test on the object's point extents — and the ring-skipping road-level search see many near
    misses), one batch, every scan against the oracle: placement choices and keep-masks exact, xyz within 1e-6 m."""
    cases = [synth.make_case(task, 930 + i, shape=GOLDEN_SHAPE, counts=counts, obj_range=(4.0, 16.0), n_cars=8 + i % 8)
             for i in range(16)]
    if task == "ss":                                   # one engine = one sequence map: give every scan the first pose / map
        for c in cases[1:]:
            c.pose, c.map_data = cases[0].pose, cases[0].map_data
    eng = Real3DEngine(task, cases[0].config, cases[0].db, max_scans=len(cases), max_points=max(len(c.pcl5) for c in cases),
                       map_data=cases[0].map_data, sub_batches=3, staged_rounds=staged)
    res = eng.augment_batch([scan_input_from_case(c) for c in cases])
    eng.close()
    placed = 0
    for case, got in zip(cases, res):
        ref, want = oracle_run(case)
        assert_matches_oracle(case, got, ref, want)
        placed += len(got.inserted)
    assert placed >= 24


def test_long_boxes_in_crowded_scenes_vs_oracle():
    """Trucks (7 m x 2.5 m boxes: 16 x 16 cells of the obstacle grid under their reach circle) among 10 - 15 parked cars:
    the collision walk clips every row of cells to the oriented box (warp_visit_clipped, od/fs:109-127 still decided by
    the exact cut_bounding_box test), so a near miss at a box corner and a hit just inside a face must both come out as
    in the oracle.  12 scans, three trucks and one motorcycle each; placement choices and keep-masks exact."""
    cases = [synth.make_case("ss", 1230 + i, shape=GOLDEN_SHAPE, counts=[0, 1, 3, 0, 0, 0], obj_range=(4.0, 16.0),
                             n_cars=10 + i % 6) for i in range(12)]
    for c in cases[1:]:
        c.pose, c.map_data = cases[0].pose, cases[0].map_data
    eng = Real3DEngine("ss", cases[0].config, cases[0].db, max_scans=len(cases), max_points=max(len(c.pcl5) for c in cases),
                       map_data=cases[0].map_data)
    res = eng.augment_batch([scan_input_from_case(c) for c in cases])
    stats = eng.stats()
    eng.close()
    placed = 0
    for case, got in zip(cases, res):
        ref, want = oracle_run(case)
        assert_matches_oracle(case, got, ref, want)
        placed += sum(1 for _, _, cls in got.inserted if int(cls) == 18)
    assert placed >= 30 and stats["walker_detail"]["n_collide"] > 100          # trucks were placed, after many collision tests


@pytest.mark.parametrize("task,seed,counts", [("od", 811, [2, 2]), ("ss", 812, [1, 1, 1, 0, 1, 0])])
def test_points_next_to_bin_edges_take_the_exact_path(task, seed, counts):
    """The ingest and the occlusion counts bin with a float estimate of the angles where that is provably safe and fall
    back to the reference's fp64 expression next to a bin edge (r3d_common.cuh, fast binning).  Scene points are moved
    onto / next to azimuth-bin edges (offsets from 1e-9 to 1e-2 of a bin, both sides, after the float32 cast of the
    scan): keep-masks, placement choices and counts must still equal the oracle's."""
    case = synth.make_case(task, seed, shape=GOLDEN_SHAPE, counts=counts, n_cars=5, obj_range=(4.0, 16.0))
    rng = np.random.default_rng(seed)
    cols = 1440
    idx = rng.choice(len(case.pcl5), size=len(case.pcl5) // 3, replace=False)
    rho = np.hypot(case.pcl5[idx, 0], case.pcl5[idx, 1])
    edge = rng.integers(0, cols + 1, len(idx))
    off = rng.choice([0.0, 1e-9, 1e-7, 1e-5, 1e-4, 1e-3, 2e-3, 3e-3, 5e-3, 1e-2], len(idx)) * rng.choice([-1.0, 1.0], len(idx))
    az = (edge + off) * (2 * np.pi / cols) - np.pi                         # od/ins:76: azimuth = arctan2(y, x) + pi
    case.pcl5[idx, 0] = (rho * np.cos(az)).astype(np.float32)
    case.pcl5[idx, 1] = (rho * np.sin(az)).astype(np.float32)
    ref, want = oracle_run(case)
    eng = make_engine(case)
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    eng.close()
    assert got.status == 0
    assert_matches_oracle(case, got, ref, want)


@pytest.mark.parametrize("variant", ["no_tma", "odd_width"])
def test_close_fill_fallback_paths_match_oracle(variant, monkeypatch):
    """The batch-wide close / fill has three tile loaders: the TMA tensor map (default), `cp.async` (R3D_NO_TMA=1, or a
    z-buffer the tensor map cannot describe) and plain loads for odd image widths (rows not 16-byte aligned).  The two
    fallbacks must give the oracle's result too."""
    case = synth.make_case("od", 821, shape=GOLDEN_SHAPE, counts=[2, 1], n_cars=5, obj_range=(4.0, 16.0))
    kw = {}
    if variant == "no_tma":
        monkeypatch.setenv("R3D_NO_TMA", "1")
    else:
        kw = dict(rows=61, cols=1439)
    eng = make_engine(case, **kw)
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    eng.close()
    ref, want = oracle_run(case, **({"num_row": 61, "num_column": 1439} if variant == "odd_width" else {}))
    assert got.status == 0
    assert_matches_oracle(case, got, ref, want)


def test_walk_profile_reports_every_scan():
    """r3d_engine_walk_profile: cycles > 0 and the number of tried cut objects per scan of the last run."""
    cases = [synth.make_case("od", 830 + i, shape=GOLDEN_SHAPE, counts=[1, 1], n_cars=4, obj_range=(4.0, 16.0)) for i in range(3)]
    eng = make_engine(cases[0], n_scans=3, max_points=max(len(c.pcl5) for c in cases))
    res = eng.augment_batch([scan_input_from_case(c) for c in cases])
    cycles, tries = eng.walk_profile()
    eng.close()
    assert len(cycles) == 3 and (cycles > 0).all()
    assert all(t >= len(r.inserted) for t, r in zip(tries, res))


def test_ground_labels_of_addjust_map_2_come_from_the_config():
    """ss/ins:209 reads a list of ground labels (`ROAD_INDEXES`) that the reference's semseg script never defines; the
    engine takes od/fs:14's [40, 44, 48] unless the config names others (`insertion.road_indexes`, as the shipped
    waymo.yaml does).  With another list the occupied map cells — and with them the placements — change, on the device
    exactly as in the oracle."""
    import copy
    case = synth.make_case("ss", 871, shape=GOLDEN_SHAPE, counts=[1, 1, 0, 1, 1, 0], n_cars=6, obj_range=(4.0, 16.0))
    ref0, _ = oracle_run(case)
    case.config = copy.deepcopy(case.config)
    case.config["insertion"]["road_indexes"] = [40]          # sidewalk / parking points now occupy their map cells
    ref, want = oracle_run(case)
    eng = make_engine(case)
    assert eng.road_indexes == [40]
    got = eng.augment_batch([scan_input_from_case(case)])[0]
    eng.close()
    assert_matches_oracle(case, got, ref, want)
    assert [(n, int(r)) for n, r, _ in ref["inserted"]] != [(n, int(r)) for n, r, _ in ref0["inserted"]], "the list must matter on this case"
