"""GPU tier: the CUDA primitives (called through the C ABI via pcl_augmentation_b200.ops) against the numpy oracle
and the golden fixtures produced by the unmodified reference."""
import numpy as np
import pytest

from oracle import real3d_oracle as orc
from pcl_augmentation_b200 import ops, synth
from tests.helpers import load_golden
from tests.test_oracle_golden import FN_IMG, _fn_projection_full_inputs, _fn_projection_inputs, check_projection_full

pytestmark = pytest.mark.gpu


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.int64)
    return np.abs(a - b)


def edge_distance_rows(el, mn, mx, rows):
    d = (mx - mn) / rows
    v = (el - mn - 0.00001) / d
    return np.abs(v - np.round(v)) * d


def test_fill_spherical_and_projection_vs_golden(monkeypatch):
    g = load_golden("fn_projection")
    pcl5, obj = _fn_projection_inputs()
    monkeypatch.setattr(ops, "NUMCOLUMN", FN_IMG[1])
    pc = ops.add_space_for_spherical(pcl5)
    pc, mx, mn = ops.fill_spherical(pc)
    np.testing.assert_array_equal(pc[:, 3], g["sph"][:, 0])                 # r: +,*,sqrt only -> bit exact
    # azimuth = atan2 + pi: libdevice vs numpy differ by <= 2 ulp of pi (absolute), whatever the size of the sum
    assert np.abs(pc[:, 4] - g["sph"][:, 1]).max() <= 1.5e-15
    assert ulp_diff(pc[:, 5], g["sph"][:, 2]).max() <= 4                    # acos
    assert abs(mx - float(g["max_el"])) < 1e-14 and abs(mn - float(g["min_el"])) < 1e-14
    train, label, pc = ops.geometrical_front_view(pc, *FN_IMG, mx, mn)
    mism = np.nonzero(pc[:, 8].astype(np.int32) != g["pix"])[0]
    # any bin flip must sit within a few ulp of a bin edge (none expected on this fixture)
    assert len(mism) == 0, (len(mism), edge_distance_rows(pc[mism, 5], mn, mx, FN_IMG[0]))
    np.testing.assert_array_equal(train, g["train"])
    np.testing.assert_array_equal(label.astype(np.int8), g["label"])
    s_train, s_label = ops.smooth_out(train, label)
    np.testing.assert_array_equal(s_label.astype(np.int8), g["s_label"])
    np.testing.assert_array_equal(s_train, g["s_train"])
    # object with the scene's elevation range, sample=True
    opc = ops.add_space_for_spherical(obj)
    opc, _, _ = ops.fill_spherical(opc)
    o_train, o_label, opc = ops.geometrical_front_view(opc, *FN_IMG, mx, mn, sample=True)
    np.testing.assert_array_equal(opc[:, 8].astype(np.int32), g["o_pix"])
    np.testing.assert_array_equal(o_train, g["o_train"])
    os_train, os_label = ops.smooth_out(o_train, o_label)
    np.testing.assert_array_equal(os_train, g["os_train"])
    np.testing.assert_array_equal(os_label.astype(np.int8), g["os_label"])


def test_projection_at_baseline_size_vs_reference_golden(monkeypatch):
    """120 000 points on the reference's 112 x 1440 image: pix ids, range image, closing and fp64 hole-fill means
    bit-identical to what the unmodified reference produced (tests/golden/fn_projection_full.npz)."""
    g = load_golden("fn_projection_full")
    pcl5 = _fn_projection_full_inputs()
    monkeypatch.setattr(ops, "NUMCOLUMN", 1440)
    pc = ops.add_space_for_spherical(pcl5)
    pc, mx, mn = ops.fill_spherical(pc)
    assert abs(mx - float(g["max_el"])) < 1e-14 and abs(mn - float(g["min_el"])) < 1e-14
    train, label, pc = ops.geometrical_front_view(pc, 112, 1440, mx, mn)
    s_train, s_label = ops.smooth_out(train, label)
    check_projection_full(g, pc, mx, mn, train, label, s_train, s_label, exact_elevation=False)


def test_projection_assert_like_reference():
    pcl5, _ = _fn_projection_inputs()
    pc = ops.add_space_for_spherical(pcl5[:2000])
    pc, mx, mn = ops.fill_spherical(pc)
    with pytest.raises(AssertionError):
        ops.geometrical_front_view(pc, 64, 512, mx - 0.05, mn + 0.05)        # rows fall outside -> od/ins:111


@pytest.mark.parametrize("shape", [(112, 1440), (64, 2048), (128, 2048), (7, 5), (1, 1), (3, 200)])
def test_close_fill_vs_oracle_random(shape):
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    for density in (0.02, 0.3, 0.7):
        label = np.where(rng.random(shape) < density, 1.0, -1.0)
        train = np.where(label == 1, rng.uniform(1, 80, shape), 500.0)
        want_t, want_l = orc.smooth_out(train, label)
        got_t, got_l = ops.smooth_out(train, label)
        np.testing.assert_array_equal(got_l, want_l)
        np.testing.assert_array_equal(got_t, want_t)
        np.testing.assert_array_equal(ops.class_closing(label), orc.class_closing(label))


def test_close_fill_full_size_scan_image():
    pcl, labels = synth.make_scan(8)
    pcl5 = np.hstack((pcl, labels.reshape(-1, 1))).astype(np.float64)
    pc = orc.add_space_for_spherical(pcl5)
    pc, mx, mn = orc.fill_spherical(pc)
    train, label, pc = orc.geometrical_front_view(pc, 112, 1440, mx, mn)
    gt, gl = ops.smooth_out(train, label)
    wt, wl = orc.smooth_out(train, label)
    np.testing.assert_array_equal(gt, wt)
    np.testing.assert_array_equal(gl, wl)


def test_projection_full_size_vs_oracle():
    pcl, labels = synth.make_scan(9)
    pcl5 = np.hstack((pcl, labels.reshape(-1, 1))).astype(np.float64)
    want = orc.add_space_for_spherical(pcl5)
    want, wmx, wmn = orc.fill_spherical(want)
    wt, wl, want = orc.geometrical_front_view(want, 112, 1440, wmx, wmn)
    got = ops.add_space_for_spherical(pcl5)
    got, mx, mn = ops.fill_spherical(got)
    np.testing.assert_array_equal(got[:, 3], want[:, 3])
    # feed the oracle's elevation range so that only the per-point transcendental ulps can matter
    gt, gl, got = ops.geometrical_front_view(got, 112, 1440, wmx, wmn)
    mism = np.nonzero(got[:, 8] != want[:, 8])[0]
    assert len(mism) == 0, (len(mism), edge_distance_rows(got[mism, 5], wmn, wmx, 112))
    np.testing.assert_array_equal(gt, wt)
    np.testing.assert_array_equal(gl, wl)


def test_empty_and_tiny_inputs():
    t, l, pc = ops.geometrical_front_view(np.zeros((0, 9)), 16, 32, 2.0, 1.0)
    assert (t == 500).all() and (l == -1).all() and pc.shape == (0, 9)
    assert ops.cut_bounding_box(np.zeros((0, 5)), {"center": {"x": 0, "y": 0, "z": 0},
                                                   "rotation": {"x": 0, "y": 0, "z": 0, "w": 1},
                                                   "length": 1, "width": 1, "height": 1, "class": "Car"}).shape == (0, 5)


def test_cut_bounding_box_vs_golden():
    g = load_golden("fn_cut_bbox")
    rng = np.random.default_rng(99)
    pts = rng.uniform(-6, 6, (6000, 5))
    pts[:, 2] = rng.uniform(-2, 3, 6000)
    pts = pts.astype(np.float32).astype(np.float64)
    for b, packed in zip(g["boxes"], g["masks"]):
        anno = {"center": {"x": b[0], "y": b[1], "z": b[2]}, "rotation": {"x": b[3], "y": b[4], "z": b[5], "w": b[6]},
                "length": b[7], "width": b[8], "height": b[9], "class": "Car"}
        m = ops.cut_bounding_box_mask(pts, anno)
        np.testing.assert_array_equal(m, np.unpackbits(packed)[:len(pts)].astype(bool))
        np.testing.assert_array_equal(ops.cut_bounding_box(pts, anno), pts[m])


def test_cut_bounding_box_boundary_points_strict():
    anno = {"center": {"x": 1.0, "y": 2.0, "z": -1.0}, "rotation": {"x": 0, "y": 0, "z": 0, "w": 1},
            "length": 2.0, "width": 1.0, "height": 1.5, "class": "Car"}
    pts = np.array([[2.0, 2.0, 0.0, 0, 0],      # on the +length face -> outside (strict <)
                    [1.999999, 2.0, 0.0, 0, 0],
                    [1.0, 2.0, -1.0, 0, 0],     # on the bottom -> outside (strict >)
                    [1.0, 2.0, 0.5, 0, 0],      # on the top -> outside
                    [1.0, 2.4999, 0.4999, 0, 0]], dtype=np.float64)
    np.testing.assert_array_equal(ops.cut_bounding_box_mask(pts, anno), orc.cut_bounding_box_mask(pts, anno))
    np.testing.assert_array_equal(ops.cut_bounding_box_mask(pts, anno), [False, True, False, False, True])
