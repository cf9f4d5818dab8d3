// Kernels of the batched Real3D-Aug engine, part: A5 - A10 placement search (on-map test, road level, collision).
// Included by r3d_engine_kernels.cuh (inside namespace r3d, after the shared constants); not a standalone header.
// ------------------------------------------------------------------------------------------------ placement
// find_possible_places (od/fs:227-304, ss/fs:192-273) for the cut object each scan is trying, as staged kernels:
//   k_onmap      one CTA per scan, one THREAD per yaw candidate, object points in shared memory: A5 + A6a (OD) with a
//                few-point prefilter, ordered list of the on-map rotations
//   k_road_level A7 for every listed rotation of EVERY scan: the (scan, rotation) tasks of the whole batch are
//                flattened (prefix sum of the per-scan list lengths) and dealt round-robin to 8-lane groups of a
//                persistent grid, so the load is balanced over all SMs whatever the per-scan distribution is
//   k_onmap_ss   (semseg) A6b with the reference's carried z shift, ordered, + the list for the next stage
//   k_collide    A8 + A9 for the listed rotations that have a road level, balanced like k_road_level
// Each task is a short chain of dependent loads (cell ranges -> a few dozen points), so throughput comes from many
// tasks in flight; 8 lanes share one task so that a dense cell near the sensor does not serialise on one thread.
constexpr unsigned CF_COLLIDE = 4u;
constexpr unsigned CF_PRE = 8u;          // internal: survived the on-map prefilter
constexpr int TRY_THREADS = 512;
constexpr int OBJ_SMEM_PTS = 1024;       // object points staged in shared memory (x, y, z fp64 = 24 KB); larger: global
constexpr int ONMAP_PRE_PTS = 8;         // points of the on-map prefilter
constexpr int GRP = 8;                   // lanes per (scan, rotation) task in the balanced stages
#ifndef R3D_TASK_THREADS
#define R3D_TASK_THREADS 256
#endif
constexpr int TASK_THREADS = R3D_TASK_THREADS;
#ifndef R3D_TASK_CTAS_PER_SM
#define R3D_TASK_CTAS_PER_SM 4
#endif
constexpr int TASK_CTAS_PER_SM = R3D_TASK_CTAS_PER_SM;

__host__ __device__ __forceinline__ size_t onmap_smem_bytes(int K) {
    return (size_t)(3 * OBJ_SMEM_PTS) * 8 + (size_t)((K + 4) & ~3) * 2 + (size_t)((K + 8) & ~7);
}

__device__ __forceinline__ bool surface_label_s(const ClassCfg& cc, unsigned lab) {
    bool ok = false;
    for (int i = 0; i < cc.n_surface; ++i) ok |= lab == (unsigned)cc.surface[i];
    return ok;
}

// up to 8 surface labels in registers (unused slots never match: labels are 16-bit)
struct SurfaceSet { unsigned l[R3D_MAX_SURFACE]; };
__device__ __forceinline__ SurfaceSet load_surface(const ClassCfg& cc) {
    SurfaceSet s;
#pragma unroll
    for (int i = 0; i < R3D_MAX_SURFACE; ++i) s.l[i] = i < cc.n_surface ? (unsigned)cc.surface[i] : 0xFFFFFFFFu;
    return s;
}
__device__ __forceinline__ bool in_surface(const SurfaceSet& s, unsigned lab) {
    bool ok = false;
#pragma unroll
    for (int i = 0; i < R3D_MAX_SURFACE; ++i) ok |= lab == s.l[i];
    return ok;
}

template <int NL = GRP>           // lane mask of the calling thread's NL-lane group (NL = 8 or 32)
__device__ __forceinline__ unsigned group_mask() {
    return NL == 32 ? 0xFFFFFFFFu : ((1u << (NL & 31)) - 1u) << ((threadIdx.x & 31) & ~(NL - 1));
}

// Visit every point stored in the grid cells [x0, x1] x [y0, y1] MINUS the cells of the hole [hx0, hx1] x [hy0, hy1]
// (a rectangle inside the first one; hx1 < hx0 = no hole) with the 8 lanes of a group.  The lanes fetch the CSR
// ranges of up to 8 rows at once (a row of cells is contiguous; a row crossing the hole has a left and a right
// segment), then stride over each segment with four independent 16-byte loads in flight per lane: the walk is a
// chain of L2 latencies, so memory-level parallelism is what counts.  `f(v)` is called per point; `stop()` is polled
// after every row (group-uniform early exit).
struct CellRect { int x0, x1, y0, y1; };
// A CSR grid to walk: the scan's whole grid in global memory, or a window of it staged in shared memory by the walker
// (cell table of the window's rows with one extra leading entry per row = the row's begin; offsets index the staged
// points).  Cell coordinates are the global ones in both cases.
struct GridView {
    const int* cell;
    const float4* pts;
    int stride, ybase, bias;        // entry of cell (x, y) = cell[(y - ybase) * stride + bias + x]
    CellRect win;                   // cells the view holds
};
__device__ __forceinline__ GridView global_view(const int* cell, const float4* pts, int G) {
    return GridView{cell, pts, G, 0, 0, CellRect{0, G - 1, 0, G - 1}};
}
__device__ __forceinline__ bool view_holds(const GridView& v, const CellRect& rc) {
    return rc.x0 >= v.win.x0 && rc.x1 <= v.win.x1 && rc.y0 >= v.win.y0 && rc.y1 <= v.win.y1;
}
template <int NL = GRP, bool STAGED = false, class F, class S>
__device__ __forceinline__ void group_visit(const GridView& gv, CellRect rc, CellRect hole, int gl, unsigned gm, F f, S stop) {
    const int* __restrict__ cell = gv.cell;
    const float4* __restrict__ pts = gv.pts;
    auto ldc = [&](int i) { return STAGED ? cell[i] : __ldg(&cell[i]); };
    for (int yb = rc.y0; yb <= rc.y1; yb += NL) {
        const int nrows = min(NL, rc.y1 - yb + 1);
        int beg_a = 0, end_a = 0, beg_b = 0, end_b = 0;
        if (gl < nrows) {
            const int y = yb + gl, row = (y - gv.ybase) * gv.stride + gv.bias;
            const bool split = hole.x1 >= hole.x0 && y >= hole.y0 && y <= hole.y1;
            const int xa1 = split ? hole.x0 - 1 : rc.x1;                 // segment A: [x0, xa1], B: [hole.x1 + 1, x1]
            if (xa1 >= rc.x0) {
                const int c0 = row + rc.x0;
                beg_a = (STAGED || c0 > 0) ? ldc(c0 - 1) : 0;
                end_a = ldc(row + xa1);
            }
            if (split && hole.x1 < rc.x1) {
                beg_b = ldc(row + hole.x1);
                end_b = ldc(row + rc.x1);
            }
        }
        // rows are dealt to sub-groups of 8 lanes (one for the 8-lane groups of the staged kernels, four for a warp): the
        // point loads of NSUB rows are in flight together instead of one dependent round trip per row
        constexpr int SL = 8, NSUB = NL / SL;
        const int sub = gl / SL, sl = gl % SL;
        for (int r0 = 0; r0 < nrows; r0 += NSUB) {
            const int r = r0 + sub;
            const bool valid = r < nrows;
#pragma unroll
            for (int seg = 0; seg < 2; ++seg) {
                const int rb = __shfl_sync(gm, seg ? beg_b : beg_a, valid ? r : 0, NL);
                int re = __shfl_sync(gm, seg ? end_b : end_a, valid ? r : 0, NL);
                if (!valid) re = rb;
                for (int p = rb + sl; p < re; p += 4 * SL) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int q = p + u * SL; if (q < re) v[u] = STAGED ? pts[q] : __ldg(&pts[q]); }
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (p + u * SL < re) f(v[u]);
                }
            }
            if (stop()) return;
        }
    }
}
// the same walk on the staged window when it holds the whole rectangle, else on the global grid
template <int NL = GRP, class F, class S>
__device__ __forceinline__ void group_visit_any(const GridView& global, const GridView* staged, CellRect rc, CellRect hole, int gl,
                                                unsigned gm, F f, S stop) {
    if (staged != nullptr && view_holds(*staged, rc)) group_visit<NL, true>(*staged, rc, hole, gl, gm, f, stop);
    else group_visit<NL, false>(global, rc, hole, gl, gm, f, stop);
}

// grid cells that hold every point within distance R of (cx, cy): grid_coord is monotone and the float conversions
// are padded by 2 mm (a float ulp at 100 m is 8 um), so no such point can sit in a cell outside
__device__ __forceinline__ CellRect cells_within(const EngineDev& e, double cx, double cy, double R) {
    const float m = 2e-3f;
    CellRect rc;
    rc.x0 = grid_coord(e, (float)(cx - R) - m); rc.x1 = grid_coord(e, (float)(cx + R) + m);
    rc.y0 = grid_coord(e, (float)(cy - R) - m); rc.y1 = grid_coord(e, (float)(cy + R) + m);
    return rc;
}

// A7 (od/fs:138-172, ss/fs:107-152): road level under a candidate centre (cx, cy), one 8-lane group.
// "first radius 0.1, 0.2, ... whose disc holds a surface point" == radius index of the NEAREST surface point, found
// by scanning the grid cells around the centre in growing square rings (each ring only visits the cells the smaller
// squares did not cover); the level is the mean z of the points inside that disc, summed in 2^-40 fixed point (order
// independent; exact for float32 z, so equal to numpy's sequential float64 sum).
template <int NL = GRP>
__device__ bool group_road_level(const EngineDev& e, int b, const SurfaceSet& surf, double cx, double cy, int gl, unsigned gm,
                                 double& level, const GridView* staged = nullptr) {
    const int G = e.G;
    const GridView gv = global_view(e.gcell + (size_t)b * G * G, e.gpts + (size_t)b * e.max_points, G);
    double best = 1e300;
    CellRect hole{0, -1, 0, -1};
    const double step = e.grid_cell;
    double R = 1.5 * step;          // first square: the cells within 0.75 m (most candidates find their nearest surface point there)
    // every cell closer than `ring` cells (Chebyshev) to the centre's cell is empty: start at the radius whose square
    // of cells still lies inside that empty block and treat the block as already visited
    const int ring = e.gnear[(size_t)b * G * G + (size_t)grid_coord(e, (float)cy) * G + grid_coord(e, (float)cx)];
    if (ring >= 2) {
        R = fmin((ring - 1) * step, 5.0);
        hole = cells_within(e, cx, cy, R - 0.01);
    }
    for (;; R = fmin(R < step ? step : R + step, 5.0)) {
        const CellRect rc = cells_within(e, cx, cy, R);
        group_visit_any<NL>(gv, staged, rc, hole, gl, gm, [&](const float4& v) {
            if (!in_surface(surf, __float_as_uint(v.w))) return;
            const double dx = sub((double)v.x, cx), dy = sub((double)v.y, cy);
            best = fmin(best, add(mul(dx, dx), mul(dy, dy)));                           // od/fs:153
        }, [] { return false; });
        for (int o = NL / 2; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(gm, best, o));
        if (best <= R * R || R >= 5.0) break;          // every point outside the scanned cells is farther than R
        hole = rc;
    }
    if (!(best <= e.radii_sq[R3D_NUM_RADII - 1])) return false;
    const int j = radius_index(e.radii_sq, best);       // smallest j with best <= r_j^2
    if (!e.radii_ok[j]) return false;                   // od/fs:156-160: no surface within reach
    const double r2 = e.radii_sq[j];
    long long zsum = 0;
    int cnt = 0;
    group_visit_any<NL>(gv, staged, cells_within(e, cx, cy, sqrt(r2)), CellRect{0, -1, 0, -1}, gl, gm, [&](const float4& v) {
        if (!in_surface(surf, __float_as_uint(v.w))) return;
        const double dx = sub((double)v.x, cx), dy = sub((double)v.y, cy);
        if (add(mul(dx, dx), mul(dy, dy)) <= r2) { zsum += __double2ll_rn(mul((double)v.z, kFix)); ++cnt; }
    }, [] { return false; });
    for (int o = NL / 2; o > 0; o >>= 1) {
        zsum += __shfl_xor_sync(gm, zsum, o);
        cnt += __shfl_xor_sync(gm, cnt, o);
    }
    if (cnt == 0) return false;
    level = __ddiv_rn(__ddiv_rn((double)zsum, kFix), (double)cnt);    // od/fs:164 np.mean
    return true;
}

// ---- lean warp-wide grid walks for the per-scan walker (64 registers per thread: the generic group_visit keeps its
// four float4 loads per lane in local memory there, which serialises them).  The four 8-lane sub-groups of the warp take
// rows y0 + sub, y0 + sub + 4, ...; every sub-group fetches its own row's CSR range (no shuffles, no hole logic: a ring
// search re-visits the inner cells, which costs two loads per row), two 16-byte point loads in flight per lane.
// `part` of `nparts`: several warps can share one rectangle, each taking every nparts-th block of four rows
template <class F, class S>
__device__ __forceinline__ void warp_visit(const int* __restrict__ cell, const float4* __restrict__ pts, int G, const CellRect& rc,
                                           int lane, F f, S stop, int part = 0, int nparts = 1) {
    const int sub = lane >> 3, sl = lane & 7;
    for (int y0 = rc.y0 + 4 * part; y0 <= rc.y1; y0 += 4 * nparts) {   // four rows of cells in flight, then the (warp-uniform) exit test
        const int y = y0 + sub;
        if (y <= rc.y1) {
            const int c0 = y * G + rc.x0;
            const int rb = c0 > 0 ? __ldg(&cell[c0 - 1]) : 0, re = __ldg(&cell[y * G + rc.x1]);
            for (int p = rb + sl; p < re; p += 16) {
                const float4 a = __ldg(&pts[p]);
                const bool two = p + 8 < re;
                float4 bq = a;
                if (two) bq = __ldg(&pts[p + 8]);
                f(a);
                if (two) f(bq);
            }
        }
        if (stop()) return;
    }
}

// ---- the same walk for the collision test of a long box: every row of cells is clipped to the columns the oriented box
// can reach.  The box is the intersection of two slabs lo < ax x + ay y < hi (YawTest); over the y range of a row of
// cells each slab bounds x by an interval that is linear in y, so its extremes sit at the row's edges.  A 10 m x 2.5 m
// truck turned by 45 degrees covers ~180 of the 441 cells of its reach square.  `clip` = 6 floats per warp in shared memory
// (A, B, S per slab: x in (A - S y, B - S y); a slab that is nearly parallel to the x axis does not bound x: A = -1e6,
// B = 1e6, S = 0); float arithmetic with 2 cm of padding on both axes (coordinates < 200 m, |S| <= 20: errors < 1e-3 m),
// so a point inside the box is never skipped; the exact test in `f` still decides.  Points beyond the grid extent are
// clamped into the border cells by grid_coord, which is monotone: border ROWS are not clipped, border columns need no care.
#ifndef R3D_CLIP_MIN_COLS
#define R3D_CLIP_MIN_COLS 4          // rectangles narrower than this many cells take the plain walk; 0 = never clip
#endif
constexpr float CLIP_PAD = 0.02f;
struct YawTest;
__device__ __forceinline__ void make_row_clip(const YawTest& t, float* clip);
template <class F, class S>
__device__ __forceinline__ void warp_visit_clipped(const EngineDev& e, const int* __restrict__ cell, const float4* __restrict__ pts,
                                                   const CellRect& rc, int lane, const float* clip, F f, S stop, int part = 0,
                                                   int nparts = 1) {
    const int sub = lane >> 3, sl = lane & 7, G = e.G;
    const float cellf = (float)e.grid_cell;
    for (int y0 = rc.y0 + 4 * part; y0 <= rc.y1; y0 += 4 * nparts) {
        const int y = y0 + sub;
        if (y <= rc.y1) {
            int x0 = rc.x0, x1 = rc.x1;
            if (y > 0 && y < G - 1) {
                const float ylo = (float)(y - (G >> 1)) * cellf - CLIP_PAD, yhi = ylo + cellf + 2.f * CLIP_PAD;
                const float a0 = clip[2] * ylo, b0 = clip[2] * yhi, a1 = clip[5] * ylo, b1 = clip[5] * yhi;
                const float xmin = fmaxf(clip[0] - fmaxf(a0, b0), clip[3] - fmaxf(a1, b1)) - CLIP_PAD;
                const float xmax = fminf(clip[1] - fminf(a0, b0), clip[4] - fminf(a1, b1)) + CLIP_PAD;
                if (xmin > xmax) x1 = x0 - 1;
                else { x0 = max(x0, grid_coord(e, xmin)); x1 = min(x1, grid_coord(e, xmax)); }
            }
            if (x0 <= x1) {
                const int c0 = y * G + x0;
                const int rb = c0 > 0 ? __ldg(&cell[c0 - 1]) : 0, re = __ldg(&cell[y * G + x1]);
                for (int p = rb + sl; p < re; p += 16) {
                    const float4 a = __ldg(&pts[p]);
                    const bool two = p + 8 < re;
                    float4 bq = a;
                    if (two) bq = __ldg(&pts[p + 8]);
                    f(a);
                    if (two) f(bq);
                }
            }
        }
        if (stop()) return;
    }
}

// A7 for one candidate centre, one warp (same result as group_road_level: the nearest surface point decides the radius
// index whatever sequence of growing squares finds it).  CHECK_LABEL = false when the surface grid holds one label only (OD).
template <bool CHECK_LABEL>
__device__ bool warp_road_level(const EngineDev& e, int b, const unsigned* __restrict__ surf, int n_surf, double cx, double cy, int lane,
                                double& level) {
    const int G = e.G;
    const int* __restrict__ cell = e.gcell + (size_t)b * G * G;
    const float4* __restrict__ pts = e.gpts + (size_t)b * e.max_points;
    auto on_surface = [&](unsigned lab) {
        if (!CHECK_LABEL) return true;
        bool ok = false;
        for (int i = 0; i < n_surf; ++i) ok |= lab == surf[i];
        return ok;
    };
#ifndef R3D_LEVEL_R0
#define R3D_LEVEL_R0 1.5                            // first search radius in cells (any value gives the same result)
#endif
    const double step = e.grid_cell;
    double R = R3D_LEVEL_R0 * step;
    const int ring = e.gnear[(size_t)b * G * G + (size_t)grid_coord(e, (float)cy) * G + grid_coord(e, (float)cx)];
    if (ring >= 3) R = fmin((ring - 1) * step, 5.0);       // every cell closer than `ring` cells is empty
    double best = 1e300;
    for (;; R = fmin(R + step, 5.0)) {
        const CellRect rc = cells_within(e, cx, cy, R);
        warp_visit(cell, pts, G, rc, lane, [&](const float4& v) {
            if (!on_surface(__float_as_uint(v.w))) return;
            const double dx = sub((double)v.x, cx), dy = sub((double)v.y, cy);
            best = fmin(best, add(mul(dx, dx), mul(dy, dy)));                           // od/fs:153
        }, [] { return false; });
        for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (best <= R * R || R >= 5.0) break;              // every point outside the scanned cells is farther than R
    }
    if (!(best <= e.radii_sq[R3D_NUM_RADII - 1])) return false;
    const int j = radius_index(e.radii_sq, best);          // smallest j with best <= r_j^2
    if (!e.radii_ok[j]) return false;                      // od/fs:156-160: no surface within reach
    const double r2 = e.radii_sq[j];
    long long zsum = 0;
    int cnt = 0;
    warp_visit(cell, pts, G, cells_within(e, cx, cy, sqrt(r2)), lane, [&](const float4& v) {
        if (!on_surface(__float_as_uint(v.w))) return;
        const double dx = sub((double)v.x, cx), dy = sub((double)v.y, cy);
        if (add(mul(dx, dx), mul(dy, dy)) <= r2) { zsum += __double2ll_rn(mul((double)v.z, kFix)); ++cnt; }
    }, [] { return false; });
    for (int o = 16; o > 0; o >>= 1) {
        zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (cnt == 0) return false;
    level = __ddiv_rn(__ddiv_rn((double)zsum, kFix), (double)cnt);    // od/fs:164 np.mean
    return true;
}

// A5 + A6a (od/fs:263-279), one thread, object points [i0, i1): every in-map point of yaw candidate k must sit on a
// map cell == 1 (returns at the first one that does not, od/fs:277-279)
struct OdMap { const unsigned char* map; int sx, sy; double mx, my; };
__device__ __forceinline__ void thread_onmap_od(const OdMap& m, const double* ox, const double* oy, int i0, int i1, double c,
                                                double sn, bool& any_in, bool& bad) {
    for (int i = i0; i < i1; ++i) {
        const double x = ox[i], y = oy[i];
        const double gx = sub(sub(mul(c, x), mul(sn, y)), m.mx);
        const double gy = sub(add(mul(sn, x), mul(c, y)), m.my);
        if (!(gx < 0.0 || gx >= (double)m.sx || gy < 0.0 || gy >= (double)m.sy)) {
            any_in = true;
            if (__ldg(&m.map[(size_t)((int)gx) * m.sy + (int)gy]) != 1) { bad = true; return; }
        }
    }
}

__device__ __forceinline__ bool obstacle_point(const EngineDev& e, int b, const ScanState& s, const ClassCfg& cc, size_t base,
                                               int p) {
    if (!e.alive[base + p]) return false;
    const unsigned lab = e.label[base + p];
    if (e.task == 0) { if (lab == (unsigned)e.road_label) return false; }         // od/ins:353-355 + od/fs:121
    else if (surface_label_s(cc, lab)) return false;                              // ss/fs:92-93
    if (s.dirty && pix_removed(e, b, s, e.pix[base + p])) return false;            // od/ins:472,491
    return true;
}

// cut_bounding_box thresholds (cb:30-66) of a YAW-ONLY box: with m = [[m00, m01, 0], [m10, m00, 0], [0, 0, 1]] the
// terms of make_box_test / inside_box that multiply a zero matrix entry are exact zeros, so dropping them changes no
// comparison: a0 = m00 x + m10 y, a1 = m01 x + m00 y, a2 = z, hi2 = cz + H, lo2 = cz.
struct YawTest { double c0x, c0y, hi0, lo0, c1x, hi1, lo1, hi2, lo2; };
__device__ __forceinline__ YawTest make_yaw_test(const YawBox& yb, double cz, double L, double W, double H) {
    YawTest t;
    t.c0x = yb.m00; t.c0y = yb.m10; t.c1x = yb.m01;
    const double hx = __ddiv_rn(mul(t.c0x, L), 2.0), hy = __ddiv_rn(mul(t.c0y, L), 2.0);
    t.hi0 = add(mul(t.c0x, add(yb.cx, hx)), mul(t.c0y, add(yb.cy, hy)));
    t.lo0 = add(mul(t.c0x, sub(yb.cx, hx)), mul(t.c0y, sub(yb.cy, hy)));
    const double wx = __ddiv_rn(mul(t.c1x, W), 2.0), wy = __ddiv_rn(mul(t.c0x, W), 2.0);
    t.hi1 = add(mul(t.c1x, add(yb.cx, wx)), mul(t.c0x, add(yb.cy, wy)));
    t.lo1 = add(mul(t.c1x, sub(yb.cx, wx)), mul(t.c0x, sub(yb.cy, wy)));
    t.hi2 = add(cz, H); t.lo2 = cz;
    return t;
}
__device__ __forceinline__ bool inside_yaw(const YawTest& t, double x, double y, double z) {
    const double a0 = add(mul(t.c0x, x), mul(t.c0y, y));
    if (!(a0 < t.hi0) || !(a0 > t.lo0)) return false;
    const double a1 = add(mul(t.c1x, x), mul(t.c0x, y));
    if (!(a1 < t.hi1) || !(a1 > t.lo1)) return false;
    return (z < t.hi2) && (z > t.lo2);
}

// column bounds of a row of grid cells under this box (see warp_visit_clipped)
__device__ __forceinline__ void make_row_clip(const YawTest& t, float* clip) {
    const double ax[2] = {t.c0x, t.c1x}, ay[2] = {t.c0y, t.c0x}, lo[2] = {t.lo0, t.lo1}, hi[2] = {t.hi0, t.hi1};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float A = -1e6f, B = 1e6f, S = 0.f;
        if (fabs(ax[k]) >= 0.05) {
            const double inv = 1.0 / ax[k], p = lo[k] * inv, q = hi[k] * inv;
            A = (float)fmin(p, q); B = (float)fmax(p, q); S = (float)(ay[k] * inv);
        }
        clip[3 * k] = A; clip[3 * k + 1] = B; clip[3 * k + 2] = S;
    }
}
// rare paths of the collision test, kept out of line so the common path stays small (per-lane results):
// obstacle points among the points of one already inserted object (its tail slice)
template <int NL = GRP>
__device__ __noinline__ bool tail_hits_candidate(const EngineDev& e, int b, const ScanState& s, const ClassCfg& cc,
                                                 const YawTest& yt, double zmin_ped, int t0, int cnt, int gl) {
    const size_t base = (size_t)b * e.P;
    const bool ped = cc.pedestrian != 0;
    for (int i = gl; i < cnt; i += NL) {
        const size_t t = (size_t)b * e.max_inserted + t0 + i;
        const double x = e.tail_x[t], y = e.tail_y[t], z = e.tail_z[t];
        if ((!ped || z >= zmin_ped) && inside_yaw(yt, x, y, z) && obstacle_point(e, b, s, cc, base, s.n0 + t0 + i))
            return true;
    }
    return false;
}
// any object point of the candidate strictly inside a scene box (od/fs:129-134)
template <int NL = GRP>
__device__ __noinline__ bool object_in_scene_box(const BoxTest* box_test, const double* ox, const double* oy, const double* oz,
                                                 int count, double c, double sn, double dz, int gl) {
    const BoxTest sbt = *box_test;
    for (int i = gl; i < count; i += 2 * NL) {
        const int i1 = i + NL;
        const double x0 = ox[i], y0 = oy[i], z0 = oz[i];
        const double x1 = i1 < count ? ox[i1] : x0, y1 = i1 < count ? oy[i1] : y0, z1 = i1 < count ? oz[i1] : z0;
        if (inside_box(sbt, sub(mul(c, x0), mul(sn, y0)), add(mul(sn, x0), mul(c, y0)), add(z0, dz))) return true;
        if (inside_box(sbt, sub(mul(c, x1), mul(sn, y1)), add(mul(sn, x1), mul(c, y1)), add(z1, dz))) return true;
    }
    return false;
}

// Can the rectangle that holds every point of candidate k (ObjBox extents in the frame of the rotated box) and a
// yaw-only scene box overlap?  Separating-axis test on the four edge directions plus the z intervals, with 1e-9 m of
// slack; "false" proves that no object point is inside the scene box, "true" only means: test the points.
__device__ __forceinline__ bool extent_may_touch_box(const ObjBox& ob, const YawBox& yb, double level, const Box& bx) {
    if (bx.m[2] != 0.0 || bx.m[5] != 0.0 || bx.m[6] != 0.0 || bx.m[7] != 0.0 || !(bx.m[8] > 0.999999)) return true;   // tilted box: no pruning
    if (level + ob.ez1 <= bx.cz - 1e-9) return false;                        // every object point at or below the box bottom
    if (level + ob.ez0 >= bx.cz + bx.height + 1e-9) return false;            // ... at or above its top
    const double ux = yb.m00, uy = yb.m10, vx = -yb.m10, vy = yb.m00;              // axes of the candidate's box
    const double s0x = bx.m[0], s0y = bx.m[3], s1x = bx.m[1], s1y = bx.m[4];        // axes of the scene box
    const double mu = 0.5 * (ob.eu0 + ob.eu1), mv = 0.5 * (ob.ev0 + ob.ev1);
    const double hu = 0.5 * (ob.eu1 - ob.eu0), hv = 0.5 * (ob.ev1 - ob.ev0);
    const double dx = yb.cx + ux * mu + vx * mv - bx.cx, dy = yb.cy + uy * mu + vy * mv - bx.cy;
    const double hl = 0.5 * bx.length, hw = 0.5 * bx.width;
    const double ax[4] = {ux, vx, s0x, s1x}, ay[4] = {uy, vy, s0y, s1y};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double nn = sqrt(ax[i] * ax[i] + ay[i] * ay[i]);
        const double ro = hu * fabs(ux * ax[i] + uy * ay[i]) + hv * fabs(vx * ax[i] + vy * ay[i]);
        const double rb = hl * fabs(s0x * ax[i] + s0y * ay[i]) + hw * fabs(s1x * ax[i] + s1y * ay[i]);
        if (fabs(dx * ax[i] + dy * ay[i]) > ro + rb + 1e-9 * nn) return false;
    }
    return true;
}

// A8 + A9 (od/fs:109-135, ss/fs:79-104) for one candidate with road level `level`, one 8-lane group:
//  (i)  obstacle scene points strictly inside the candidate box: the ORIGINAL points come from the all-points grid
//       (only the cells within the box reach of the candidate centre), the INSERTED points from the tails of the
//       already placed objects whose box is close enough;
//  (ii) object points strictly inside an existing / already inserted box.
// The exact cut_bounding_box test (strict inequalities in the reference's expression order) decides; grid cells and
// bounding circles only prune.
// (part 2 of the collision test) obstacle points among the inserted objects' tails + object points inside scene boxes
template <int NL = GRP>
__device__ __noinline__ bool group_collides_rest(const EngineDev& e, int b, const ScanState& s, const ObjBox& ob, const ClassCfg& cc,
                                                 const YawBox& yb, const YawTest& yt, double c, double sn, double level, int gl,
                                                 unsigned gm);

template <int NL = GRP>
__device__ bool group_collides(const EngineDev& e, int b, const ScanState& s, const ObjBox& ob, const ClassCfg& cc, double c,
                               double sn, double level, int gl, unsigned gm, const GridView* staged = nullptr) {
    const YawBox yb = make_yaw_box(ob.cx, ob.cy, ob.a, ob.b, c, sn);
    const YawTest yt = make_yaw_test(yb, level, ob.length, ob.width, ob.height);
    const double zmin_ped = add(level, 0.1);                                      // od/fs:123-124
    const bool ped = cc.pedestrian != 0;
    const size_t base = (size_t)b * e.P;
    bool hit = false;
    {
        const int G = e.G;
        const float fcx = (float)yb.cx, fcy = (float)yb.cy, fr = (float)ob.reach + 1e-3f, fr2 = fr * fr;
        const float zlo = (float)level - 1e-3f, zhi = (float)(level + ob.height) + 1e-3f;
        const CellRect rc{grid_coord(e, fcx - fr), grid_coord(e, fcx + fr), grid_coord(e, fcy - fr), grid_coord(e, fcy + fr)};
        const unsigned ok_slots = e.task == 1 ? cc.ok_slots : 0u;       // OD tags carry slot 0 and no class accepts it here
        group_visit_any<NL>(global_view(e.acell + (size_t)b * G * G, e.apts + (size_t)b * e.max_points, G), staged, rc,
                            CellRect{0, -1, 0, -1}, gl, gm, [&](const float4& v) {
            if (hit) return;
            const unsigned tag = __float_as_uint(v.w);
            if ((ok_slots >> (tag >> APT_IDX_BITS)) & 1u) return;        // semseg: ground the class may stand on (ss/fs:92-93)
            const float dx = v.x - fcx, dy = v.y - fcy;                  // cheap conservative pruning first
            if (dx * dx + dy * dy > fr2 || v.z < zlo || v.z > zhi) return;
            const double x = v.x, yy = v.y, z = v.z;
            if (ped && !(z >= zmin_ped)) return;
            if (!inside_yaw(yt, x, yy, z)) return;                       // exact test (cb:30-66)
            if (obstacle_point(e, b, s, cc, base, (int)(tag & APT_IDX_MASK))) hit = true;
        }, [&] { return (__ballot_sync(gm, hit) & gm) != 0u; });
        if (__ballot_sync(gm, hit) & gm) return true;
    }
    return group_collides_rest<NL>(e, b, s, ob, cc, yb, yt, c, sn, level, gl, gm);
}

template <int NL>
__device__ __noinline__ bool group_collides_rest(const EngineDev& e, int b, const ScanState& s, const ObjBox& ob, const ClassCfg& cc,
                                                 const YawBox& yb, const YawTest& yt, double c, double sn, double level, int gl,
                                                 unsigned gm) {
    const double zmin_ped = add(level, 0.1);
    const int gshift = (threadIdx.x & 31) & ~(NL - 1);
    bool hit = false;
    // the lanes look at 8 boxes at a time; only the boxes whose bounding circle reaches the candidate's are tested
    const int nbox0 = s.n_boxes - s.n_inserted;
    int t_run = 0;
    for (int j0 = 0; j0 < s.n_inserted; j0 += NL) {                     // the tail of placed object j lies inside its box
        const int j = j0 + gl;
        int cnt = 0;
        bool near = false;
        if (j < s.n_inserted) {
            cnt = e.inserted[((size_t)b * e.max_events + j) * 4 + 3];
            const Box& bx = e.boxes[(size_t)b * e.max_boxes + nbox0 + j];
            const double ddx = yb.cx - bx.cx, ddy = yb.cy - bx.cy, rr = ob.reach + bx.reach + 0.05;
            near = ddx * ddx + ddy * ddy <= rr * rr;
        }
        int inc = cnt;                                                   // tail offsets: prefix sum of the point counts
        for (int o = 1; o < NL; o <<= 1) { const int t = __shfl_up_sync(gm, inc, o, NL); if (gl >= o) inc += t; }
        unsigned m = (__ballot_sync(gm, near) & gm) >> gshift;
        while (m) {
            const int q = __ffs(m) - 1; m &= m - 1;
            const int qcnt = __shfl_sync(gm, cnt, q, NL), qt0 = t_run + __shfl_sync(gm, inc, q, NL) - qcnt;
            hit = tail_hits_candidate<NL>(e, b, s, cc, yt, zmin_ped, qt0, qcnt, gl);
            if (__ballot_sync(gm, hit) & gm) return true;
        }
        t_run += __shfl_sync(gm, inc, NL - 1, NL);
    }
    const double dz = sub(level, ob.cz);
    const double *ox = e.obj_x + ob.first, *oy = e.obj_y + ob.first, *oz = e.obj_z + ob.first;
    for (int b0 = 0; b0 < s.n_boxes; b0 += NL) {                        // (ii) od/fs:129-134
        const int bi = b0 + gl;
        bool near = false;
        if (bi < s.n_boxes) {
            const Box& bx = e.boxes[(size_t)b * e.max_boxes + bi];
            const double ddx = yb.cx - bx.cx, ddy = yb.cy - bx.cy, rr = ob.reach + bx.reach + 0.05;
            near = ddx * ddx + ddy * ddy <= rr * rr && extent_may_touch_box(ob, yb, level, bx);
        }
        unsigned m = (__ballot_sync(gm, near) & gm) >> gshift;
        while (m) {
            const int q = __ffs(m) - 1; m &= m - 1;
            hit = object_in_scene_box<NL>(&e.box_tests[(size_t)b * e.max_boxes + b0 + q], ox, oy, oz, ob.count, c, sn, dz, gl);
            if (__ballot_sync(gm, hit) & gm) return true;
        }
    }
    return false;
}

// A flag / hint word in shared memory that OTHER warps set while this one polls it (the shared hit flag of the warps that
// split a candidate, the collision witness): read and written with shared-memory atomics, lane 0 reads for the warp
// (any value is acceptable at any time, the atomics only make the accesses race-free for the memory model / racecheck)
__device__ __forceinline__ int warp_peek(const volatile int* p, int lane) {
    int v = 0;
    if (lane == 0) v = atomicAdd(const_cast<int*>(p), 0);
    return __shfl_sync(0xffffffffu, v, 0);
}
// A8 + A9 for one candidate, one warp: part (i) on the lean walk, the rest shared with group_collides
// `part` of `nparts` warps work on the same candidate (the walker hands spare warps of a chunk to its candidates): the
// rows of the obstacle grid are dealt to the parts, the tails / scene boxes are the last part's; `shared_hit` (shared
// memory, zeroed by the caller) carries a hit to the other parts
// `witness` (shared memory, -1 = none): index of the original point that made an earlier candidate of this try collide.
// Neighbouring yaws turn the box by a fraction of a degree, so the same obstacle usually sits in the next candidate's box
// too: it is tested first (the same exact test as in the walk below, so a hit is a hit), and most colliding candidates of a
// try that fails never walk the grid at all.
__device__ bool warp_collides(const EngineDev& e, int b, const ScanState& s, const ObjBox& ob, const ClassCfg& cc, double c, double sn,
                              double level, int lane, int part = 0, int nparts = 1, volatile int* shared_hit = nullptr,
                              int* witness = nullptr, float* clip = nullptr) {
    const YawBox yb = make_yaw_box(ob.cx, ob.cy, ob.a, ob.b, c, sn);
    const YawTest yt = make_yaw_test(yb, level, ob.length, ob.width, ob.height);
    {
        const double zmin_ped = add(level, 0.1);                                  // od/fs:123-124
        const bool ped = cc.pedestrian != 0;
        const size_t base = (size_t)b * e.P;
        if (witness != nullptr) {
            const int wp = warp_peek(witness, lane);
            if (wp >= 0) {
                const float4 v = __ldg(&e.xyzi[(size_t)b * e.max_points + wp]);
                const double x = v.x, yy = v.y, z = v.z;
                if ((!ped || z >= zmin_ped) && inside_yaw(yt, x, yy, z) && obstacle_point(e, b, s, cc, base, wp)) return true;
            }
        }
        const int G = e.G;
        const float fcx = (float)yb.cx, fcy = (float)yb.cy, fr = (float)ob.reach + 1e-3f, fr2 = fr * fr;
        const float zlo = (float)level - 1e-3f, zhi = (float)(level + ob.height) + 1e-3f;
        const CellRect rc{grid_coord(e, fcx - fr), grid_coord(e, fcx + fr), grid_coord(e, fcy - fr), grid_coord(e, fcy + fr)};
        const unsigned ok_slots = e.task == 1 ? cc.ok_slots : 0u;       // OD tags carry slot 0 and no class accepts it here
        bool hit = false;
        int hit_p = -1;
        auto test_point = [&](const float4& v) {
            if (hit) return;
            const unsigned tag = __float_as_uint(v.w);
            if ((ok_slots >> (tag >> APT_IDX_BITS)) & 1u) return;        // semseg: ground the class may stand on (ss/fs:92-93)
            const float dx = v.x - fcx, dy = v.y - fcy;                  // cheap conservative pruning first
            if (dx * dx + dy * dy > fr2 || v.z < zlo || v.z > zhi) return;
            const double x = v.x, yy = v.y, z = v.z;
            if (ped && !(z >= zmin_ped)) return;
            if (!inside_yaw(yt, x, yy, z)) return;                       // exact test (cb:30-66)
            if (obstacle_point(e, b, s, cc, base, (int)(tag & APT_IDX_MASK))) { hit = true; hit_p = (int)(tag & APT_IDX_MASK); }
        };
        auto stop = [&] { return __any_sync(0xffffffffu, hit) != 0 || (shared_hit != nullptr && warp_peek(shared_hit, lane) != 0); };
        if (R3D_CLIP_MIN_COLS > 0 && clip != nullptr && rc.x1 - rc.x0 + 1 >= R3D_CLIP_MIN_COLS) {     // warp-uniform
            if (lane == 0) make_row_clip(yt, clip);
            __syncwarp();
            warp_visit_clipped(e, e.acell + (size_t)b * G * G, e.apts + (size_t)b * e.max_points, rc, lane, clip, test_point, stop,
                               part, nparts);
            __syncwarp();                                                  // the next candidate of this warp rewrites `clip`
        } else {
            warp_visit(e.acell + (size_t)b * G * G, e.apts + (size_t)b * e.max_points, G, rc, lane, test_point, stop, part, nparts);
        }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm) {
            if (witness != nullptr && lane == __ffs(hm) - 1) atomicExch(witness, hit_p);
            return true;
        }
    }
    if (part != nparts - 1 || (shared_hit != nullptr && warp_peek(shared_hit, lane) != 0)) return false;
    return group_collides_rest<32>(e, b, s, ob, cc, yb, yt, c, sn, level, lane, 0xffffffffu);
}

// ordered compaction of the rotations 1..K whose flag byte satisfies (f & mask) == want (all threads of the CTA);
// ends with a barrier
__device__ int block_compact(const unsigned char* flags, int K, unsigned mask, unsigned want, unsigned short* list, int* s_warp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int base = 0;
    for (int k0 = 1; k0 <= K; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        const bool ok = k <= K && (flags[k] & mask) == want;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_warp[w] = __popc(m);
        __syncthreads();
        int off = base, tot = 0;
        for (int i = 0; i < nw; ++i) { if (i < w) off += s_warp[i]; tot += s_warp[i]; }
        if (ok) list[off + __popc(m & ((1u << lane) - 1u))] = (unsigned short)k;
        base += tot;
        __syncthreads();
    }
    return base;
}

// stage 1: per-scan set-up of the try (object record, cleared candidate arrays) + OD on-map test
__global__ void __launch_bounds__(TRY_THREADS) k_onmap(EngineDev e, int n_scans) {
    const int b = blockIdx.x;
    if (b >= n_scans || !e.gate_try[b]) { if (b < n_scans && threadIdx.x == 0) e.n_list[b] = 0; return; }
    extern __shared__ double s_place[];
    __shared__ int s_warp[TRY_THREADS / 32];
    const int K = e.K, tid = threadIdx.x;
    double* s_ox = s_place;
    double* s_oy = s_ox + OBJ_SMEM_PTS;
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_oy + 2 * OBJ_SMEM_PTS);      // [K + 1]
    unsigned char* s_flags = reinterpret_cast<unsigned char*>(s_list + ((K + 4) & ~3));       // [K + 1]
    const ScanState& s = e.st[b];
    const ObjBox ob = e.obj[s.cur_obj];
    const size_t cb = (size_t)b * (K + 1);
    if (tid == 0) e.try_obj[b] = ob;
    for (int k = tid; k <= K; k += TRY_THREADS) { e.cand_level[cb + k] = 0.0; e.cand_v[cb + k] = 0; }
    if (e.task == 1) {                   // semseg: the road level comes first, for every rotation (ss/fs:229)
        for (int k = tid; k <= K; k += TRY_THREADS) e.cand_flags[cb + k] = 0;
        if (tid == 0) e.n_list[b] = K;
        return;
    }
    const double *ox = e.obj_x + ob.first, *oy = e.obj_y + ob.first;
    if (ob.count <= OBJ_SMEM_PTS) {
        for (int i = tid; i < ob.count; i += TRY_THREADS) { s_ox[i] = ox[i]; s_oy[i] = oy[i]; }
        ox = s_ox; oy = s_oy;
    }
    for (int k = tid; k <= K; k += TRY_THREADS) s_flags[k] = 0;
    OdMap m;
    {
        const int msel = e.classes[ob.cls].map_sel;
        const int* dims = e.od_map_dims + ((size_t)b * 2 + msel) * 4;
        m.sx = dims[0]; m.sy = dims[1]; m.mx = (double)dims[2]; m.my = (double)dims[3];
        m.map = e.od_maps + e.od_map_off[(size_t)b * 2 + msel];
    }
    __syncthreads();
    // the first few points decide most rotations (off the road)
    const int npre = min(ob.count, ONMAP_PRE_PTS);
    for (int k = 1 + tid; k <= K; k += TRY_THREADS) {
        bool any_in = false, bad = false;
        thread_onmap_od(m, ox, oy, 0, npre, e.cos_k[k], e.sin_k[k], any_in, bad);
        if (!bad) s_flags[k] = CF_PRE;
    }
    __syncthreads();
    const int n = block_compact(s_flags, K, CF_PRE, CF_PRE, e.cand_list + cb, s_warp);
    for (int k = tid; k <= K; k += TRY_THREADS) e.cand_flags[cb + k] = s_flags[k];
    if (tid == 0) { e.n_list[b] = n; atomicAdd(&e.stats[6], (unsigned long long)n); }
}

// prefix sum of the per-scan list lengths into shared memory (every CTA of a balanced stage); returns the total
// phase 0: every listed rotation; 1: the first cand_window of each scan; 2: the rest, for the scans flagged in need2
__device__ __forceinline__ int task_count(const EngineDev& e, int b, int phase) {
    if (!e.gate_try[b]) return 0;
    const int n = e.n_list[b];
    if (phase == 0) return n;
    if (phase == 1) return min(n, e.cand_window);
    return e.need2[b] ? max(n - e.cand_window, 0) : 0;
}
__device__ int task_prefix(const EngineDev& e, int n_scans, int* s_pref, int phase = 0) {
    __shared__ int s_w[TASK_THREADS / 32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) { s_carry = 0; s_pref[0] = 0; }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < n_scans; b0 += TASK_THREADS) {
        const int b = b0 + threadIdx.x;
        const int v = b < n_scans ? task_count(e, b, phase) : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        int off = s_carry;
        for (int i = 0; i < w; ++i) off += s_w[i];
        if (b < n_scans) s_pref[b + 1] = off + inc;
        __syncthreads();
        if (threadIdx.x == TASK_THREADS - 1) s_carry = off + inc;
    }
    __syncthreads();
    return s_pref[n_scans];
}
__device__ __forceinline__ int task_scan(const int* s_pref, int n_scans, int t) {      // largest b with s_pref[b] <= t
    int lo = 0, hi = n_scans - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_pref[mid] <= t) lo = mid; else hi = mid - 1; }
    return lo;
}

// stage 1b (OD): A5 + A6a (od/fs:263-279) on ALL object points for the rotations that survived the prefilter; the
// 8 lanes of a group stride over the points, two per lane in flight, early-out on the first off-road point
__global__ void __launch_bounds__(TASK_THREADS, TASK_CTAS_PER_SM) k_onmap_full(const __grid_constant__ EngineDev e, int n_scans,
                                                                                int phase) {
    extern __shared__ int s_pref[];
    const int total = task_prefix(e, n_scans, s_pref, phase);
    const int first = phase == 2 ? e.cand_window : 0;
    const int gl = threadIdx.x & (GRP - 1);
    const unsigned gm = group_mask();
    const int n_groups = gridDim.x * (TASK_THREADS / GRP);
    int n_on = 0;
    for (int t = blockIdx.x * (TASK_THREADS / GRP) + threadIdx.x / GRP; t < total; t += n_groups) {
        const int b = task_scan(s_pref, n_scans, t), i = t - s_pref[b];
        const size_t cb = (size_t)b * (e.K + 1);
        const int k = e.cand_list[cb + first + i];
        const ObjBox& ob = e.try_obj[b];
        const int first = ob.first, count = ob.count, msel = e.classes[ob.cls].map_sel;
        const int* dims = e.od_map_dims + ((size_t)b * 2 + msel) * 4;
        const int sx = dims[0], sy = dims[1];
        const double mx = (double)dims[2], my = (double)dims[3];
        const unsigned char* map = e.od_maps + e.od_map_off[(size_t)b * 2 + msel];
        const double c = e.cos_k[k], sn = e.sin_k[k];
        const double *ox = e.obj_x + first, *oy = e.obj_y + first;
        bool any_in = false, bad = false;
        for (int i0 = 0; i0 < count; i0 += 4 * GRP) {
            double x[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = min(i0 + u * GRP + gl, count - 1);              // the clamped repeats change nothing
                x[u] = ox[i]; y[u] = oy[i];
            }
            unsigned char v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = 1;
                const double gx = sub(sub(mul(c, x[u]), mul(sn, y[u])), mx), gy = sub(add(mul(sn, x[u]), mul(c, y[u])), my);
                if (!(gx < 0.0 || gx >= (double)sx || gy < 0.0 || gy >= (double)sy)) {
                    any_in = true; v[u] = __ldg(&map[(size_t)((int)gx) * sy + (int)gy]);
                }
            }
            bad = v[0] != 1 || v[1] != 1 || v[2] != 1 || v[3] != 1;
            if (__ballot_sync(gm, bad) & gm) { bad = true; break; }       // od/fs:277-279
        }
        const bool on = (__ballot_sync(gm, any_in) & gm) != 0u && !bad;
        if (gl == 0) { e.cand_flags[cb + k] = on ? CF_ONMAP : 0; n_on += on; }
    }
    for (int o = 16; o > 0; o >>= 1) n_on += __shfl_xor_sync(0xffffffffu, n_on, o);
    if ((threadIdx.x & 31) == 0 && n_on) atomicAdd(&e.stats[7], (unsigned long long)n_on);
}

// stage 2: A7 for the listed rotations (OD: the on-map ones, od/fs:281; semseg: all, the map test comes after)
__global__ void __launch_bounds__(TASK_THREADS, TASK_CTAS_PER_SM) k_road_level(const __grid_constant__ EngineDev e, int n_scans,
                                                                                int phase) {
    extern __shared__ int s_pref[];
    const int total = task_prefix(e, n_scans, s_pref, phase);
    const int first = phase == 2 ? e.cand_window : 0;
    const int gl = threadIdx.x & (GRP - 1);
    const unsigned gm = group_mask();
    const int n_groups = gridDim.x * (TASK_THREADS / GRP);
    for (int t = blockIdx.x * (TASK_THREADS / GRP) + threadIdx.x / GRP; t < total; t += n_groups) {
        const int b = task_scan(s_pref, n_scans, t), i = t - s_pref[b];
        const size_t cb = (size_t)b * (e.K + 1);
        const int k = e.task == 0 ? (int)e.cand_list[cb + first + i] : i + 1;
        if (e.task == 0 && !(e.cand_flags[cb + k] & CF_ONMAP)) continue;         // failed the full on-map test
        const ObjBox& ob = e.try_obj[b];
        const SurfaceSet surf = load_surface(e.classes[ob.cls]);
        const double c = e.cos_k[k], sn = e.sin_k[k], ocx = ob.cx, ocy = ob.cy;
        double level = 0.0;
        const bool ok = group_road_level(e, b, surf, sub(mul(c, ocx), mul(sn, ocy)), add(mul(sn, ocx), mul(c, ocy)), gl, gm, level);
        if (gl == 0 && ok) { e.cand_flags[cb + k] = (e.task == 0 ? CF_ONMAP : 0u) | CF_HOK; e.cand_level[cb + k] = level; }
    }
}

// stage 2b (semseg): A6b (ss/fs:231-248) with the reference's carried z shift (ss/fs:146-147 is in place):
// world = T . [x y z 1] - move, astype(int); every in-map cell value must be allowed.  The reference visits the yaws in
// order because the z shift of the last yaw that passed AND found a road level feeds the test of the next one:
//   pass_k = f(k, dz_k),   dz_{k+1} = pass_k && level_k ? level_k - cz : dz_k,   dz_1 = 0.
// k_onmap_ss solves that recurrence by fixed-point iteration: every yaw is tested in parallel (one warp per yaw) under
// an ASSUMED dz, one thread then walks the pass flags in order and derives the dz every yaw should have seen, and the
// yaws whose assumption was wrong are tested again.  After i sweeps the first i yaws are final (induction), so the
// loop ends with the sequential solution; in practice the shift moves a point by millimetres (it enters the world xy
// only through the tilt of the pose) and two or three sweeps suffice.  k_onmap_ss_seq is the literal ordered walk,
// kept for more than SS_MAX_K yaws.
constexpr int SS_MAX_K = 1024;

struct SsMapTest {          // everything the map test of one object point needs
    double t00, t01, t02, t03, t10, t11, t12, t13;
    unsigned okmask;
    const unsigned* occ;
    const int* far;         // occupied cells outside the bit window (numpy-wrapped indices of addjust_map_2), [0] = count
    double inv02, inv12;    // 1 / |t02|, 1 / |t12| (1e300 for a zero term): cell-edge distance -> tolerated change of the z shift
};
// `tol` (when asked for): how far the carried z shift may move from `dz` before this point can reach another map cell.
// dz enters the cell indices only through t02 * z and t12 * z (the pitch / roll terms of the pose), so a point that is
// `dx` away from the nearest integer of (wx - move_x) keeps its cell while |shift| * |t02| < dx (1e-9 m covers the rounding
// of the expression, ~1e-13 m); same for y.  The walker retests a yaw only when the carried shift left this interval.
template <bool TOL>
__device__ __forceinline__ bool ss_point_off_map_t(const EngineDev& e, const ScanState& s, const SsMapTest& m, double x0, double y0,
                                                   double z0, double c, double sn, double dz, double& tol) {
    const double x = sub(mul(c, x0), mul(sn, y0)), y = add(mul(sn, x0), mul(c, y0));
    const double z = add(z0, dz);
    const double wx = add(add(add(mul(m.t00, x), mul(m.t01, y)), mul(m.t02, z)), m.t03);
    const double wy = add(add(add(mul(m.t10, x), mul(m.t11, y)), mul(m.t12, z)), m.t13);
    const double vx = sub(wx, (double)e.ss_move_x), vy = sub(wy, (double)e.ss_move_y);
    if (TOL) {
        const double dx = fabs(vx - rint(vx)) - 1e-9, dy = fabs(vy - rint(vy)) - 1e-9;
        tol = fmin(dx * m.inv02, dy * m.inv12);
    }
    const int ix = trunc_to_int(vx);
    const int iy = trunc_to_int(vy);
    if (ix < e.ss_sx && ix > -1 && iy < e.ss_sy && iy > -1) {
        unsigned v = e.ss_map[(size_t)ix * e.ss_sy + iy];
        const int lx = ix - s.win_x0, ly = iy - s.win_y0;
        if (lx >= 0 && ly >= 0 && lx < e.map_window && ly < e.map_window) {
            const int bit = lx * e.map_window + ly;
            if (m.occ[bit >> 5] & (1u << (bit & 31))) v = 4;
        } else if (m.far[0] > 0) {
            const int cell = ix * e.ss_sy + iy;
            for (int i = 0; i < OCC_FAR_CAP; ++i) if (m.far[1 + i] == cell) v = 4;
        }
        if (!((m.okmask >> v) & 1u)) return true;
    }
    return false;
}
__device__ __forceinline__ bool ss_point_off_map(const EngineDev& e, const ScanState& s, const SsMapTest& m, double x0, double y0,
                                                 double z0, double c, double sn, double dz) {
    double tol;
    return ss_point_off_map_t<false>(e, s, m, x0, y0, z0, c, sn, dz, tol);
}

__global__ void __launch_bounds__(1024) k_onmap_ss_seq(EngineDev e, int n_scans);

__global__ void __launch_bounds__(1024) k_onmap_ss(EngineDev e, int n_scans) {
    const int b = blockIdx.x;
    if (b >= n_scans || !e.gate_try[b]) return;
    const int K = e.K;
    const ScanState& s = e.st[b];
    const ObjBox ob = e.try_obj[b];
    const double* T = e.poses + (size_t)b * 16;
    SsMapTest m;
    m.t00 = T[0]; m.t01 = T[1]; m.t02 = T[2]; m.t03 = T[3]; m.t10 = T[4]; m.t11 = T[5]; m.t12 = T[6]; m.t13 = T[7];
    m.okmask = e.classes[ob.cls].map_ok_mask;
    m.occ = e.occ_win + (size_t)b * (e.map_window * e.map_window / 32);
    m.far = e.occ_far + (size_t)b * (OCC_FAR_CAP + 1);
    const size_t cb = (size_t)b * (K + 1);
    __shared__ double s_dz[SS_MAX_K + 1];              // the dz yaw k was (or must be) tested under
    __shared__ unsigned char s_pass[SS_MAX_K + 1], s_todo[SS_MAX_K + 1], s_hok[SS_MAX_K + 1];
    __shared__ int s_changed;
    for (int k = threadIdx.x; k <= K; k += blockDim.x) {
        s_dz[k] = 0.0; s_pass[k] = 0; s_todo[k] = 1;
        s_hok[k] = (e.cand_flags[cb + k] & CF_HOK) ? 1 : 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const double *ox = e.obj_x + ob.first, *oy = e.obj_y + ob.first, *oz = e.obj_z + ob.first;
    for (int sweep = 0; sweep <= K; ++sweep) {
        for (int k = 1 + warp; k <= K; k += n_warps) {           // one warp per yaw that needs (re)testing
            if (!s_todo[k]) continue;
            const double c = e.cos_k[k], sn = e.sin_k[k], dz = s_dz[k];
            bool bad = false;
            for (int i0 = 0; i0 < ob.count && !bad; i0 += 32) {
                const int i = i0 + lane;
                const bool off = i < ob.count && ss_point_off_map(e, s, m, ox[i], oy[i], oz[i], c, sn, dz);
                bad = __ballot_sync(0xffffffffu, off) != 0u;
            }
            if (lane == 0) s_pass[k] = bad ? 0 : 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {                                   // the ordered walk over the flags (ss/fs:144-148)
            double run = 0.0;
            int changed = 0;
            for (int k = 1; k <= K; ++k) {
                const bool redo = run != s_dz[k];                 // yaw k was tested under another shift: test it again
                s_todo[k] = redo;
                if (redo) { s_dz[k] = run; changed = 1; }
                // carry on with the flag at hand: if the retest flips it, the next walk re-derives everything behind k
                if (s_pass[k] && s_hok[k]) run = sub(e.cand_level[cb + k], ob.cz);
            }
            s_changed = changed;
        }
        __syncthreads();
        if (!s_changed) break;
    }
    for (int k = 1 + threadIdx.x; k <= K; k += blockDim.x)
        if (s_pass[k]) e.cand_flags[cb + k] = (unsigned char)(e.cand_flags[cb + k] | CF_ONMAP);
    __threadfence_block();
    __syncthreads();
    __shared__ int s_warp[32];
    const int n = block_compact(e.cand_flags + cb, K, CF_ONMAP | CF_HOK, CF_ONMAP | CF_HOK, e.cand_list + cb, s_warp);
    if (threadIdx.x == 0) e.n_list[b] = n;
}

__global__ void __launch_bounds__(1024) k_onmap_ss_seq(EngineDev e, int n_scans) {
    const int b = blockIdx.x;
    if (b >= n_scans || !e.gate_try[b]) return;
    const ScanState& s = e.st[b];
    const ObjBox ob = e.try_obj[b];
    const double* T = e.poses + (size_t)b * 16;
    SsMapTest m;
    m.t00 = T[0]; m.t01 = T[1]; m.t02 = T[2]; m.t03 = T[3]; m.t10 = T[4]; m.t11 = T[5]; m.t12 = T[6]; m.t13 = T[7];
    m.okmask = e.classes[ob.cls].map_ok_mask;
    m.occ = e.occ_win + (size_t)b * (e.map_window * e.map_window / 32);
    m.far = e.occ_far + (size_t)b * (OCC_FAR_CAP + 1);
    const size_t cb = (size_t)b * (e.K + 1);
    double dz = 0.0;
    for (int k = 1; k <= e.K; ++k) {
        const double c = e.cos_k[k], sn = e.sin_k[k];
        int bad = 0;
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x)
            if (ss_point_off_map(e, s, m, e.obj_x[ob.first + i], e.obj_y[ob.first + i], e.obj_z[ob.first + i], c, sn, dz)) bad = 1;
        bad = __syncthreads_or(bad);
        if (!bad) {
            const unsigned f = e.cand_flags[cb + k];
            if (f & CF_HOK) dz = sub(e.cand_level[cb + k], ob.cz);              // ss/fs:144-148
            __syncthreads();
            if (threadIdx.x == 0) e.cand_flags[cb + k] = (unsigned char)(f | CF_ONMAP);
        }
    }
    __threadfence_block();
    __syncthreads();
    __shared__ int s_warp[32];
    const int n = block_compact(e.cand_flags + cb, e.K, CF_ONMAP | CF_HOK, CF_ONMAP | CF_HOK, e.cand_list + cb, s_warp);
    if (threadIdx.x == 0) e.n_list[b] = n;
}

// stage 3: A8 + A9 for the listed rotations that have a road level (OD: the list still holds every on-map rotation)
__global__ void __launch_bounds__(TASK_THREADS, TASK_CTAS_PER_SM) k_collide(const __grid_constant__ EngineDev e, int n_scans,
                                                                             int phase) {
    extern __shared__ int s_pref[];
    const int total = task_prefix(e, n_scans, s_pref, phase);
    const int first = phase == 2 ? e.cand_window : 0;
    const int gl = threadIdx.x & (GRP - 1);
    const unsigned gm = group_mask();
    const int n_groups = gridDim.x * (TASK_THREADS / GRP);
    for (int t = blockIdx.x * (TASK_THREADS / GRP) + threadIdx.x / GRP; t < total; t += n_groups) {
        const int b = task_scan(s_pref, n_scans, t), i = t - s_pref[b];
        const size_t cb = (size_t)b * (e.K + 1);
        const int k = e.cand_list[cb + first + i];
        if (!(e.cand_flags[cb + k] & CF_HOK)) continue;
        const ObjBox& ob = e.try_obj[b];
        if (group_collides(e, b, e.st[b], ob, e.classes[ob.cls], e.cos_k[k], e.sin_k[k], e.cand_level[cb + k], gl, gm) && gl == 0)
            e.cand_flags[cb + k] = CF_ONMAP | CF_HOK | CF_COLLIDE;
    }
}

// ordered list of the feasible rotations only (the probe API; the engine rounds get it from k_occl_count)
__global__ void __launch_bounds__(128) k_feasible_list(EngineDev e, int n_scans) {
    const int b = blockIdx.x;
    if (b >= n_scans || !e.gate_try[b]) return;
    __shared__ int s_warp[4];
    __shared__ unsigned short s_feas[4096];
    const size_t cb = (size_t)b * (e.K + 1);
    int nf = 0;
    if (e.K <= 4096) {
        nf = block_compact(e.cand_flags + cb, e.K, CF_ONMAP | CF_HOK | CF_COLLIDE, CF_ONMAP | CF_HOK, s_feas, s_warp);
        for (int i = threadIdx.x; i < nf; i += blockDim.x) e.feas[(size_t)b * e.K + i] = s_feas[i];
    } else if (threadIdx.x == 0) {
        for (int k = 1; k <= e.K; ++k)
            if ((e.cand_flags[cb + k] & 7u) == (CF_ONMAP | CF_HOK)) e.feas[(size_t)b * e.K + nf++] = k;
    }
    if (threadIdx.x == 0) { e.st[b].n_feasible = nf; e.st[b].found_rank = INT_MAX; }
}
