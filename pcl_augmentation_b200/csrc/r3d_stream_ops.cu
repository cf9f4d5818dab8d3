// Stream-level primitives of the placement / occlusion / insertion stages on the reference's own float64 working-row
// layout: r3d_place_candidates (A5 + A6 + A7), r3d_obb_collide (A8 + A9), r3d_occlude_mask (A11),
// r3d_compact_insert (A12 / A14).  Device pointers in, caller-allocated device outputs, no hidden state, re-entrant
// per CUDA stream.  The batched engine (r3d_engine.cu) is the fast path for many scans; these calls are what the
// function-level drop-ins (check_bounding_box, the occlusion step of insertion.py) and other callers bind.
#include "r3d_common.cuh"
#include "r3d_host.h"
#include "../../include/real3d_b200.h"

using namespace r3d;

namespace {

constexpr int OPS_THREADS = 256;
constexpr int OPS_CHUNK = 4096;
constexpr double kFixOps = 1099511627776.0;      // 2^40 fixed point: order-independent exact sum of float32-exact z

__device__ __forceinline__ Box load_box(const double* p) {
    Box b;
    b.cx = p[0]; b.cy = p[1]; b.cz = p[2];
    for (int i = 0; i < 9; ++i) b.m[i] = p[3 + i];
    b.length = p[12]; b.width = p[13]; b.height = p[14]; b.reach = p[15];
    return b;
}

struct LabelSet8 { int n; int v[R3D_MAX_SURFACE]; };
__device__ __forceinline__ bool in_set(const LabelSet8& s, int lab) {
    bool ok = false;
    for (int i = 0; i < s.n; ++i) ok |= lab == s.v[i];
    return ok;
}

// ---------------------------------------------------------------------------------- A8 + A9: r3d_obb_collide
// (i) obstacle scene points strictly inside candidate k's box: CTA = (chunk of scene points, candidate)
__global__ void __launch_bounds__(OPS_THREADS) k_collide_scene(const double* __restrict__ rows9, int64_t n, const double* __restrict__ cand_boxes,
                                                               int mode, int pedestrian, LabelSet8 ok, unsigned char* __restrict__ out) {
    const int k = blockIdx.y;
    if (out[k]) return;                                         // another chunk already found a collision
    __shared__ BoxTest s_bt;
    __shared__ double s_zmin;
    if (threadIdx.x == 0) {
        const Box b = load_box(cand_boxes + (size_t)k * R3D_BOX_DOUBLES);
        s_bt = make_box_test(b);
        s_zmin = add(b.cz, 0.1);                                // od/fs:123-124
    }
    __syncthreads();
    const BoxTest bt = s_bt;
    bool hit = false;
    const int64_t p0 = (int64_t)blockIdx.x * OPS_CHUNK;
    for (int64_t p = p0 + threadIdx.x; p < min(p0 + (int64_t)OPS_CHUNK, n) && !hit; p += OPS_THREADS) {
        const double* r = rows9 + p * 9;
        const int lab = (int)r[7];
        const bool obstacle = mode == 0 ? lab == 1 : !in_set(ok, lab);            // od/fs:121, ss/fs:92-93
        if (!obstacle) continue;
        if (mode == 0 && pedestrian && !(r[2] >= s_zmin)) continue;
        hit = inside_box(bt, r[0], r[1], r[2]);
    }
    if (__syncthreads_or(hit) && threadIdx.x == 0) out[k] = 1;
}
// (ii) candidate k's object points strictly inside any scene box: one CTA per candidate, scene box tests in shared memory
__global__ void __launch_bounds__(OPS_THREADS) k_collide_object(const double* __restrict__ obj, int64_t m, int stride,
                                                                const double* __restrict__ cand4, const double* __restrict__ scene_boxes,
                                                                int n_boxes, unsigned char* __restrict__ out) {
    const int k = blockIdx.x;
    if (out[k]) return;
    extern __shared__ BoxTest s_boxes[];
    const double c = cand4[(size_t)k * 4], sn = cand4[(size_t)k * 4 + 1], dz = cand4[(size_t)k * 4 + 2];
    bool hit = false;
    for (int b0 = 0; b0 < n_boxes && !hit; b0 += 64) {
        const int nb = min(64, n_boxes - b0);
        __syncthreads();
        for (int j = threadIdx.x; j < nb; j += OPS_THREADS) s_boxes[j] = make_box_test(load_box(scene_boxes + (size_t)(b0 + j) * R3D_BOX_DOUBLES));
        __syncthreads();
        for (int64_t i = threadIdx.x; i < m && !hit; i += OPS_THREADS) {
            const double* p = obj + i * stride;
            const double x = sub(mul(c, p[0]), mul(sn, p[1])), y = add(mul(sn, p[0]), mul(c, p[1])), z = add(p[2], dz);
            for (int j = 0; j < nb; ++j) if (inside_box(s_boxes[j], x, y, z)) { hit = true; break; }   // od/fs:129-134
        }
        hit = __syncthreads_or(hit);
    }
    if (hit && threadIdx.x == 0) out[k] = 1;
}

// ---------------------------------------------------------------------------------------- A11: r3d_occlude_mask
__global__ void __launch_bounds__(OPS_THREADS) k_vis_px(const double* __restrict__ scene_smooth, const double* __restrict__ obj_smooth,
                                                        int num_pix, unsigned char* __restrict__ vis) {
    for (int i = blockIdx.x * OPS_THREADS + threadIdx.x; i < num_pix; i += gridDim.x * OPS_THREADS)
        vis[i] = obj_smooth[i] < scene_smooth[i] ? 1 : 0;                          // od/ins:486 (strict; empty = 500 both sides)
}
__global__ void __launch_bounds__(OPS_THREADS) k_keep_masks(const double* __restrict__ rows9, int64_t n, const unsigned char* __restrict__ vis,
                                                            int num_pix, int keep_if_visible, unsigned char* __restrict__ keep,
                                                            int* __restrict__ counter) {
    int cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * OPS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * OPS_THREADS) {
        const int pix = (int)rows9[i * 9 + 8];
        const bool v = pix >= 0 && pix < num_pix && vis[pix];
        const bool k = keep_if_visible ? v : !v;                                   // od/ins:488-501: scene loses, object keeps
        keep[i] = k ? 1 : 0;
        cnt += keep_if_visible ? k : !k;                                           // object points kept / scene points removed
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(counter, cnt);
}

// ------------------------------------------------------------------------------------ A12 / A14: r3d_compact_insert
// one CTA: stable compaction of the kept scene rows, then the kept object rows in (pix_id, index) order
__global__ void __launch_bounds__(1024) k_compact_insert(const double* __restrict__ scene, const unsigned char* __restrict__ skeep, int64_t n,
                                                         const double* __restrict__ obj, const unsigned char* __restrict__ okeep, int64_t m,
                                                         double* __restrict__ out, long long* __restrict__ n_out,
                                                         unsigned long long* __restrict__ keys, int key_cap) {
    __shared__ int s_w[32];
    __shared__ long long s_run;
    __shared__ int s_nk;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_run = 0; s_nk = 0; }
    __syncthreads();
    for (int64_t i0 = 0; i0 < n; i0 += 1024) {
        const int64_t i = i0 + threadIdx.x;
        const bool k = i < n && skeep[i];
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        if (lane == 0) s_w[w] = __popc(bal);
        __syncthreads();
        long long off = s_run;
        int tot = 0;
        for (int j = 0; j < 32; ++j) { if (j < w) off += s_w[j]; tot += s_w[j]; }
        if (k) {
            const double* src = scene + i * 9;
            double* dst = out + (off + __popc(bal & ((1u << lane) - 1u))) * 9;
            for (int c = 0; c < 9; ++c) dst[c] = src[c];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    const long long n_scene = s_run;
    for (int64_t i = threadIdx.x; i < m; i += 1024)
        if (okeep[i]) keys[atomicAdd(&s_nk, 1)] = ((unsigned long long)(unsigned)(int)obj[i * 9 + 8] << 32) | (unsigned)i;
    __syncthreads();
    const int nk = s_nk;
    int np2 = 1;
    while (np2 < nk) np2 <<= 1;
    for (int i = nk + threadIdx.x; i < np2 && i < key_cap; i += 1024) keys[i] = R3D_EMPTY_U64;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += 1024) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], bb = keys[ixj];
                    if ((a > bb) == ((i & k) == 0)) { keys[i] = bb; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    for (int j = threadIdx.x; j < nk; j += 1024) {
        const double* src = obj + (size_t)(keys[j] & 0xffffffffull) * 9;
        double* dst = out + (n_scene + j) * 9;
        for (int c = 0; c < 9; ++c) dst[c] = src[c];
    }
    if (threadIdx.x == 0) { n_out[0] = n_scene + nk; n_out[1] = n_scene; }
}

// --------------------------------------------------------------------------- A5 + A6 + A7: r3d_place_candidates
struct PlaceArgs {
    int task, K;
    double cx, cy, cz;                  // box centre (z = bottom) of the cut object as read from its annotation
    int sx, sy;
    long long move_x, move_y;           // OD: min_x / min_y of the map; semseg: map move
    double T[8];                        // semseg: first two rows of the lidar -> world pose
    unsigned ok_mask;                   // semseg: bit v set iff map value v is allowed (insertion.placement[class])
    LabelSet8 surface;                  // labels the road-level search accepts
    double radii_sq[R3D_NUM_RADII];
    int radii_ok[R3D_NUM_RADII];
};

// One CTA walks the yaw candidates IN ORDER (the semseg map test of candidate k sees the z shift of the last candidate
// that passed and found a road level, ss/fs:146-147); every step is parallel over the object points / the ground rows.
__global__ void __launch_bounds__(1024) k_place_candidates(const double* __restrict__ obj, int64_t m, int stride, const double* __restrict__ cos_k,
                                                           const double* __restrict__ sin_k, const unsigned char* __restrict__ map,
                                                           const double* __restrict__ ground, int64_t n_g, PlaceArgs a,
                                                           unsigned char* __restrict__ flags, double* __restrict__ level_out) {
    __shared__ unsigned long long s_best;
    __shared__ long long s_zsum;
    __shared__ int s_cnt, s_any, s_bad;
    double dz = 0.0;
    if (threadIdx.x == 0) { flags[0] = 0; level_out[0] = 0.0; }
    for (int k = 1; k <= a.K; ++k) {
        const double c = cos_k[k], sn = sin_k[k];
        if (threadIdx.x == 0) { s_best = R3D_EMPTY_U64; s_zsum = 0; s_cnt = 0; s_any = 0; s_bad = 0; }
        __syncthreads();
        bool any_in = false, bad = false;
        for (int64_t i = threadIdx.x; i < m && !bad; i += 1024) {
            const double* p = obj + i * stride;
            const double x = sub(mul(c, p[0]), mul(sn, p[1])), y = add(mul(sn, p[0]), mul(c, p[1]));
            if (a.task == 0) {                                   // od/fs:267-279
                const double gx = sub(x, (double)a.move_x), gy = sub(y, (double)a.move_y);
                if (!(gx < 0.0 || gx >= (double)a.sx || gy < 0.0 || gy >= (double)a.sy)) {
                    any_in = true;
                    bad = map[(size_t)((int)gx) * a.sy + (int)gy] != 1;
                }
            } else {                                             // ss/fs:235-248
                const double z = add(p[2], dz);
                const double wx = add(add(add(mul(a.T[0], x), mul(a.T[1], y)), mul(a.T[2], z)), a.T[3]);
                const double wy = add(add(add(mul(a.T[4], x), mul(a.T[5], y)), mul(a.T[6], z)), a.T[7]);
                const int ix = trunc_to_int(sub(wx, (double)a.move_x)), iy = trunc_to_int(sub(wy, (double)a.move_y));
                if (ix < a.sx && ix > -1 && iy < a.sy && iy > -1) bad = !((a.ok_mask >> map[(size_t)ix * a.sy + iy]) & 1u);
            }
        }
        if (any_in) s_any = 1;
        if (bad) s_bad = 1;
        __syncthreads();
        const bool on = a.task == 0 ? (s_any && !s_bad) : !s_bad;                  // OD starts False (od/fs:264), semseg True (ss/fs:232)
        unsigned f = on ? 1u : 0u;
        double level = 0.0;
        if (on) {                                                // correct_height (od/fs:138-172, ss/fs:107-152)
            const double ccx = sub(mul(c, a.cx), mul(sn, a.cy)), ccy = add(mul(sn, a.cx), mul(c, a.cy));
            unsigned long long best = R3D_EMPTY_U64;
            for (int64_t i = threadIdx.x; i < n_g; i += 1024) {
                const double* g = ground + i * 5;
                if (!in_set(a.surface, (int)g[4]) || !(g[2] > -3.0)) continue;
                const double dx = sub(g[0], ccx), dy = sub(g[1], ccy);
                best = min(best, dbl_bits(add(mul(dx, dx), mul(dy, dy))));
            }
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            if ((threadIdx.x & 31) == 0) atomicMin(&s_best, best);
            __syncthreads();
            const double b2 = bits_dbl(s_best);
            int j = R3D_NUM_RADII;
            if (s_best != R3D_EMPTY_U64 && b2 <= a.radii_sq[R3D_NUM_RADII - 1]) { j = 0; while (!(b2 <= a.radii_sq[j])) ++j; }
            if (j < R3D_NUM_RADII && a.radii_ok[j]) {
                const double r2 = a.radii_sq[j];
                long long zs = 0;
                int cnt = 0;
                for (int64_t i = threadIdx.x; i < n_g; i += 1024) {
                    const double* g = ground + i * 5;
                    if (!in_set(a.surface, (int)g[4]) || !(g[2] > -3.0)) continue;
                    const double dx = sub(g[0], ccx), dy = sub(g[1], ccy);
                    if (add(mul(dx, dx), mul(dy, dy)) <= r2) { zs += __double2ll_rn(mul(g[2], kFixOps)); ++cnt; }
                }
                for (int o = 16; o > 0; o >>= 1) { zs += __shfl_xor_sync(0xffffffffu, zs, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
                if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd((unsigned long long*)&s_zsum, (unsigned long long)zs); atomicAdd(&s_cnt, cnt); }
                __syncthreads();
                if (s_cnt > 0) {
                    level = __ddiv_rn(__ddiv_rn((double)s_zsum, kFixOps), (double)s_cnt);      // np.mean (od/fs:164)
                    f |= 2u;
                    dz = sub(level, a.cz);                       // the shift the next candidates' points carry (ss/fs:146-147)
                }
            }
        }
        if (threadIdx.x == 0) { flags[k] = (unsigned char)f; level_out[k] = level; }
        __syncthreads();
    }
}

}  // namespace

extern "C" int r3d_obb_collide(const double* scene_rows9, int64_t n, const double* scene_boxes, int32_t n_boxes, const double* obj_rows,
                               int64_t m, int32_t row_stride, const double* cand4, const double* cand_boxes, int32_t n_cand, int32_t mode,
                               int32_t pedestrian, const int32_t* ok_labels, int32_t n_ok, uint8_t* collide_out, r3d_stream stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n < 0 || m < 0 || n_cand <= 0 || n_boxes < 0 || !cand4 || !cand_boxes || !collide_out || (n > 0 && !scene_rows9) ||
        (m > 0 && !obj_rows) || (n_boxes > 0 && !scene_boxes) || row_stride < 3 || n_ok < 0 || n_ok > R3D_MAX_SURFACE || (mode != 0 && mode != 1))
        return r3d_fail(R3D_ERR_ARG, "r3d_obb_collide: bad argument");
    LabelSet8 ok;
    ok.n = n_ok;
    for (int i = 0; i < R3D_MAX_SURFACE; ++i) ok.v[i] = i < n_ok ? ok_labels[i] : -1;
    R3D_CUDA(cudaMemsetAsync(collide_out, 0, n_cand, st));
    if (n > 0) {
        k_collide_scene<<<dim3((unsigned)((n + OPS_CHUNK - 1) / OPS_CHUNK), n_cand), OPS_THREADS, 0, st>>>(scene_rows9, n, cand_boxes, mode,
                                                                                                            pedestrian, ok, collide_out);
        r3d_count_launch();
    }
    if (m > 0 && n_boxes > 0) {
        k_collide_object<<<n_cand, OPS_THREADS, 64 * sizeof(BoxTest), st>>>(obj_rows, m, row_stride, cand4, scene_boxes, n_boxes, collide_out);
        r3d_count_launch();
    }
    return r3d_check_launch("r3d_obb_collide");
}

extern "C" int r3d_occlude_mask(const double* scene_rows9, int64_t n, const double* obj_rows9, int64_t m, const double* scene_smooth,
                                const double* obj_smooth, int32_t num_pix, uint8_t* scene_keep, uint8_t* obj_keep, uint8_t* vis_px,
                                int32_t* counts_out, r3d_stream stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n < 0 || m < 0 || num_pix <= 0 || !scene_smooth || !obj_smooth || !vis_px || !counts_out || (n > 0 && (!scene_rows9 || !scene_keep)) ||
        (m > 0 && (!obj_rows9 || !obj_keep)))
        return r3d_fail(R3D_ERR_ARG, "r3d_occlude_mask: bad argument");
    R3D_CUDA(cudaMemsetAsync(counts_out, 0, 2 * sizeof(int32_t), st));
    k_vis_px<<<std::min((num_pix + OPS_THREADS - 1) / OPS_THREADS, 148 * 8), OPS_THREADS, 0, st>>>(scene_smooth, obj_smooth, num_pix, vis_px);
    r3d_count_launch();
    if (n > 0) {
        k_keep_masks<<<(int)std::min<int64_t>((n + OPS_THREADS - 1) / OPS_THREADS, 148 * 8), OPS_THREADS, 0, st>>>(scene_rows9, n, vis_px, num_pix, 0,
                                                                                                               scene_keep, counts_out);
        r3d_count_launch();
    }
    if (m > 0) {
        k_keep_masks<<<(int)std::min<int64_t>((m + OPS_THREADS - 1) / OPS_THREADS, 148 * 8), OPS_THREADS, 0, st>>>(obj_rows9, m, vis_px, num_pix, 1,
                                                                                                               obj_keep, counts_out + 1);
        r3d_count_launch();
    }
    return r3d_check_launch("r3d_occlude_mask");
}

extern "C" int r3d_compact_insert(const double* scene_rows9, const uint8_t* scene_keep, int64_t n, const double* obj_rows9, const uint8_t* obj_keep,
                                  int64_t m, double* out_rows9, int64_t* n_out, uint64_t* sort_scratch, int64_t scratch_len, r3d_stream stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    int64_t need = 1;
    while (need < m) need <<= 1;
    if (n < 0 || m < 0 || m > (1 << 24) || !out_rows9 || !n_out || (n > 0 && (!scene_rows9 || !scene_keep)) ||
        (m > 0 && (!obj_rows9 || !obj_keep || !sort_scratch || scratch_len < need)))
        return r3d_fail(R3D_ERR_ARG, "r3d_compact_insert: bad argument (sort_scratch needs next_pow2(m) entries)");
    k_compact_insert<<<1, 1024, 0, st>>>(scene_rows9, scene_keep, n, obj_rows9, obj_keep, m, out_rows9, reinterpret_cast<long long*>(n_out),
                                         reinterpret_cast<unsigned long long*>(sort_scratch), (int)need);
    r3d_count_launch();
    return r3d_check_launch("r3d_compact_insert");
}

extern "C" int r3d_place_candidates(const double* obj_rows, int64_t m, int32_t row_stride, const double* box8_host, int32_t yaw_steps,
                                    const double* cos_k, const double* sin_k, int32_t task, const uint8_t* map, int32_t size_x, int32_t size_y,
                                    int64_t move_x, int64_t move_y, const double* pose16_host, uint32_t map_ok_mask, const double* ground_rows5,
                                    int64_t n_ground, const int32_t* surface_labels, int32_t n_surface, const double* radii_sq_host,
                                    const int32_t* radii_ok_host, uint8_t* flags_out, double* level_out, r3d_stream stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (m <= 0 || !obj_rows || row_stride < 3 || !box8_host || yaw_steps <= 0 || !cos_k || !sin_k || (task != 0 && task != 1) || !map ||
        size_x <= 0 || size_y <= 0 || (task == 1 && !pose16_host) || n_ground < 0 || (n_ground > 0 && !ground_rows5) || !surface_labels ||
        n_surface <= 0 || n_surface > R3D_MAX_SURFACE || !radii_sq_host || !radii_ok_host || !flags_out || !level_out)
        return r3d_fail(R3D_ERR_ARG, "r3d_place_candidates: bad argument");
    PlaceArgs a;
    a.task = task; a.K = yaw_steps; a.cx = box8_host[0]; a.cy = box8_host[1]; a.cz = box8_host[2];
    a.sx = size_x; a.sy = size_y; a.move_x = move_x; a.move_y = move_y; a.ok_mask = map_ok_mask;
    for (int i = 0; i < 8; ++i) a.T[i] = task == 1 ? pose16_host[i] : 0.0;
    a.surface.n = n_surface;
    for (int i = 0; i < R3D_MAX_SURFACE; ++i) a.surface.v[i] = i < n_surface ? surface_labels[i] : -1;
    for (int i = 0; i < R3D_NUM_RADII; ++i) { a.radii_sq[i] = radii_sq_host[i]; a.radii_ok[i] = radii_ok_host[i]; }
    k_place_candidates<<<1, 1024, 0, st>>>(obj_rows, m, row_stride, cos_k, sin_k, map, ground_rows5, n_ground, a, flags_out, level_out);
    r3d_count_launch();
    return r3d_check_launch("r3d_place_candidates");
}
