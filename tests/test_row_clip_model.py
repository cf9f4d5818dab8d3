"""CPU model of the row clip of the collision walk (make_row_clip / warp_visit_clipped, csrc/r3d_k_placement.cuh): the
float32 column bounds of every row of grid cells must cover the cell of EVERY point strictly inside the oriented box
(od/fs:109-127: cut_bounding_box decides, the grid only prunes).  The device code itself is covered by the GPU parity
tests (the whole suite was also run with the clip forced on every collision test, profiles/r2_clip_ab.json)."""
import numpy as np

f32 = np.float32
G, CELL, INV, PAD = 240, f32(0.5), f32(2.0), f32(0.02)


def grid_coord(v):
    return max(0, min(int(np.floor(f32(v) * INV)) + G // 2, G - 1))


def make_row_clip(ax, ay, lo, hi):
    clip = []
    for k in range(2):
        a, b, s = f32(-1e6), f32(1e6), f32(0)
        if abs(ax[k]) >= 0.05:
            inv = 1.0 / ax[k]
            p, q = lo[k] * inv, hi[k] * inv
            a, b, s = f32(min(p, q)), f32(max(p, q)), f32(ay[k] * inv)
        clip += [a, b, s]
    return clip


def row_columns(clip, rc, y):
    x0, x1 = rc[0], rc[1]
    if 0 < y < G - 1:
        ylo = f32(f32(y - G // 2) * CELL - PAD)
        yhi = f32(ylo + CELL + f32(2) * PAD)
        a0, b0, a1, b1 = clip[2] * ylo, clip[2] * yhi, clip[5] * ylo, clip[5] * yhi
        xmin = f32(max(f32(clip[0] - max(a0, b0)), f32(clip[3] - max(a1, b1))) - PAD)
        xmax = f32(min(f32(clip[1] - min(a0, b0)), f32(clip[4] - min(a1, b1))) + PAD)
        if xmin > xmax:
            return x0, x0 - 1
        x0, x1 = max(x0, grid_coord(xmin)), min(x1, grid_coord(xmax))
    return x0, x1


def test_clipped_rows_cover_every_point_inside_the_box():
    rng = np.random.default_rng(7)
    inside = missed = cells_full = cells_clip = 0
    for it in range(600):
        yaw = rng.uniform(-np.pi, np.pi)
        if it % 5 == 0:                                         # boxes nearly aligned with the grid axes (|ax| < 0.05 branch)
            yaw = rng.choice([0, np.pi / 2, np.pi, -np.pi / 2]) + rng.uniform(-0.06, 0.06)
        c, s = np.cos(yaw), np.sin(yaw)
        length, width = rng.uniform(0.5, 12), rng.uniform(0.4, 3)
        cx, cy = rng.uniform(-70, 70, 2)                        # beyond the 60 m half extent too: clamped border cells
        ax, ay = [c, -s], [s, c]
        lo = [c * cx + s * cy - length / 2, -s * cx + c * cy - width / 2]
        hi = [c * cx + s * cy + length / 2, -s * cx + c * cy + width / 2]
        clip = make_row_clip(ax, ay, lo, hi)
        fr = f32(np.hypot(length, width) / 2) + f32(1e-3)
        rc = (grid_coord(f32(cx) - fr), grid_coord(f32(cx) + fr), grid_coord(f32(cy) - fr), grid_coord(f32(cy) + fr))
        rows = {y: row_columns(clip, rc, y) for y in range(rc[2], rc[3] + 1)}
        cells_full += (rc[1] - rc[0] + 1) * (rc[3] - rc[2] + 1)
        cells_clip += sum(max(0, x1 - x0 + 1) for x0, x1 in rows.values())
        a = rng.uniform(-length / 2, length / 2, 300)
        b = rng.uniform(-width / 2, width / 2, 300)
        a[:40] = np.sign(a[:40]) * (length / 2 - 1e-9)          # points on the faces and in the corners
        b[20:60] = np.sign(b[20:60]) * (width / 2 - 1e-9)
        px, py = (cx + c * a - s * b).astype(f32), (cy + s * a + c * b).astype(f32)
        for x, y in zip(px, py):
            a0, a1 = c * float(x) + s * float(y), -s * float(x) + c * float(y)
            if not (lo[0] < a0 < hi[0] and lo[1] < a1 < hi[1]):
                continue
            inside += 1
            gx, gy = grid_coord(x), grid_coord(y)
            if gy not in rows or not rows[gy][0] <= gx <= rows[gy][1]:
                missed += 1
    assert inside > 100_000 and missed == 0
    assert cells_clip < 0.5 * cells_full                        # and it does prune (2.9x fewer cells on this mix)
