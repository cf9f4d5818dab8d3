// Kernels of the batched Real3D-Aug engine.  One "round" = one tried cut object for every unfinished scan of the
// batch; the reference's nested Python loops (od/ins:375-614) are a per-scan state machine advanced on the device
// (k_ctrl), so the host only launches a fixed kernel sequence per round and polls one counter.
#pragma once
#include "r3d_engine.cuh"
#include "r3d_closefill.cuh"

namespace r3d {

constexpr int CHUNK = 4096;          // points per CTA of the streaming kernels (256 threads x 16 points)
constexpr int STREAM_THREADS = 256;

__device__ __forceinline__ void set_error(ScanState& s, int code) {
    if (s.status == 0) s.status = code;
}

// ------------------------------------------------------------------------------------------------ ingest
// A1 + A2 (od/ins:55-82) once per original point: r, elevation and the azimuth bin are cached in HBM.
__global__ void __launch_bounds__(STREAM_THREADS) k_ingest(EngineDev e, int n_scans) {
    const int b = blockIdx.y;
    if (b >= n_scans) return;
    ScanState& s = e.st[b];
    const int n0 = s.n0;
    const int p0 = blockIdx.x * CHUNK;
    if (p0 >= n0) return;
    const double d_az = kTwoPi / (double)e.cols;
    const size_t base = (size_t)b * e.P;
    for (int p = p0 + threadIdx.x; p < min(p0 + CHUNK, n0); p += STREAM_THREADS) {
        const float4 v = __ldg(&e.xyzi[(size_t)b * e.max_points + p]);
        const double x = v.x, y = v.y, z = v.z;
        const double r = range3(x, y, z);
        const double el = elevation(z, r);
        const int c = trunc_to_int(__ddiv_rn(az_mod(azimuth(x, y)), d_az));
        if (!(r > 0.0) || c < 0 || c >= e.cols) set_error(s, R3D_ERR_ASSERT);      // od/ins:113 / nan elevation
        e.r[base + p] = r;
        e.el[base + p] = el;
        e.col[base + p] = (unsigned short)max(0, min(c, e.cols - 1));
        e.alive[base + p] = 1;
    }
}

__global__ void k_reset_alive(EngineDev e, int n_scans) {
    const int b = blockIdx.y;
    if (b >= n_scans) return;
    const int n0 = e.st[b].n0;
    const size_t base = (size_t)b * e.P;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n0; p += gridDim.x * blockDim.x) e.alive[base + p] = 1;
}

// per-scan scheduling state from the pre-drawn counts (generate_seed, od/ins:171-187)
__global__ void k_reset_state(EngineDev e, int n_scans, const int* n0_arr, const int* nbox0_arr) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_scans) return;
    ScanState& s = e.st[b];
    s.n0 = n0_arr[b]; s.n_tail = 0; s.tail_before = 0; s.n_boxes = nbox0_arr[b];
    s.phase = PH_INIT; s.status = 0;
    s.inserted_class = -1;
    for (int c = 0; c < R3D_MAX_CLASSES; ++c) s.remaining[c] = c < e.n_classes ? e.counts[b * e.n_classes + c] : 0;
    for (int c = e.n_classes - 1; c >= 0; --c) if (s.remaining[c] > 0) s.inserted_class = c;
    s.timeout = 0; s.start_idx = 0; s.end_idx = 0; s.s_idx = 0; s.event = 0;
    s.cur_class = 0; s.cur_obj = -1;
    s.try_active = 0; s.need_project = 0; s.apply_flag = 0; s.dirty = 0; s.scene_changed = 1;
    s.n_feasible = 0; s.found_rank = INT_MAX; s.chosen_rot = 0; s.accepted = 0; s.chosen_v = 0;
    s.n_inserted = 0; s.n_check = 0; s.far_flag = 0;
    s.first = 1; s.extreme_removed = 0; s.d_r0 = 0; s.d_r1 = -1; s.d_c0 = 0; s.d_c1 = -1;
    s.new_min_bits = R3D_EMPTY_U64; s.new_max_bits = 0ull;
    s.min_el_bits = R3D_EMPTY_U64; s.max_el_bits = 0ull;
    if (e.task == 1) {     // semseg: window of map cells around the sensor for the occupied-cell overlay
        const double* T = e.poses + (size_t)b * 16;
        s.win_x0 = (int)(T[3] - (double)e.ss_move_x) - e.map_window / 2;
        s.win_y0 = (int)(T[7] - (double)e.ss_move_y) - e.map_window / 2;
    } else { s.win_x0 = 0; s.win_y0 = 0; }
    const int uw = (e.n_objects + 31) / 32;
    for (int w = 0; w < uw; ++w) e.unplaceable[(size_t)b * uw + w] = 0u;
}

// ------------------------------------------------------------------------------------------------- ctrl
// A13 (od/ins:386-428, 587-614): which cut object each scan tries next.  Thread 0 advances the reference's
// slot / window / try loops until the scan needs GPU work again; the CTA then re-arms the candidate arrays.
__device__ int list_entry(const EngineDev& e, int b, int ev, int ci, int sidx, int len) {
    const int* head = e.perms + (((size_t)b * e.n_perm_events + ev) * e.n_classes + ci) * e.max_tries;
    if (sidx < e.max_tries) return head[sidx];
    int nhead = 0;
    while (nhead < e.max_tries && head[nhead] >= 0) ++nhead;
    int want = sidx - nhead, seen = 0;
    for (int j = 0; j < len; ++j) {
        bool in_head = false;
        for (int h = 0; h < nhead; ++h) if (head[h] == j) { in_head = true; break; }
        if (in_head) continue;
        if (seen == want) return j;
        ++seen;
    }
    return -1;
}

__global__ void __launch_bounds__(32) k_ctrl(EngineDev e, int n_scans) {
    const int b = blockIdx.x;
    if (b >= n_scans) return;
    ScanState& s = e.st[b];
    if (threadIdx.x == 0) {
        const unsigned round = *(volatile unsigned*)&e.round_ctl[0];
        const int slot = (int)(round & 63u);
        int apply = 0, project = 0, tryact = 0;
        if (s.phase != PH_DONE && s.phase != PH_ERROR) {
            const int uw = (e.n_objects + 31) / 32;
            unsigned* unpl = e.unplaceable + (size_t)b * uw;
            int ci = s.cur_class;
            bool new_slot = false, new_window = false, next_try = false;
            if (s.phase == PH_INIT) {
                new_slot = true;
            } else {
                const int len = e.class_list_off[ci + 1] - e.class_list_off[ci];
                if (s.accepted) {                                   // od/ins:536-547
                    s.timeout = 0;
                    s.remaining[ci] -= 1;
                    apply = 1; s.dirty = 0; s.scene_changed = 1;
                    new_slot = true;
                } else {
                    unpl[s.cur_obj >> 5] |= 1u << (s.cur_obj & 31);   // od/ins:464-466, 583-585
                    if (s.n_feasible > 0) s.dirty = 1;                // od/ins:472,491: last failed candidate persists
                    const int sidx = s.s_idx;
                    if (sidx == len - 1 || sidx == 3 * e.max_tries) { s.remaining[ci] = 0; s.timeout = 1; }   // :587-591
                    if (sidx == s.end_idx - 1) {                                                             // :595-614
                        s.remaining[ci] -= 1;
                        if (s.remaining[ci] <= 0) new_slot = true; else new_window = true;
                    } else { s.s_idx = sidx + 1; next_try = true; }
                }
            }
            for (int guard = 0; guard < 100000; ++guard) {
                if (new_slot) {
                    new_slot = false;
                    int mx = 0;
                    for (int c = 0; c < e.n_classes; ++c) mx = max(mx, s.remaining[c]);
                    if (mx <= 0) {                                   // od/ins:375
                        s.phase = PH_DONE;
                        if (s.dirty) { apply = 1; s.dirty = 0; s.tail_before = s.n_tail; }
                        break;
                    }
                    if (s.dirty) { apply = 1; s.dirty = 0; s.scene_changed = 1; s.tail_before = s.n_tail; }
                    if (s.scene_changed) { project = 1; s.scene_changed = 0; }
                    for (int c = 0; c < e.n_classes; ++c)
                        if (s.remaining[c] > 0) { ci = c; break; }    // od/ins:386-391
                    if (s.inserted_class != ci) s.timeout = 0;
                    s.inserted_class = ci; s.cur_class = ci;
                    new_window = true;
                }
                const int len = e.class_list_off[ci + 1] - e.class_list_off[ci];
                if (new_window) {
                    new_window = false;
                    if (!s.timeout) {                                // od/ins:399-402 (random.shuffle = next table row)
                        if (s.event >= e.n_perm_events) { set_error(s, R3D_ERR_INDEX); s.phase = PH_ERROR; break; }
                        s.event += 1;
                        s.start_idx = 0; s.end_idx = e.max_tries;
                    } else {                                         // od/ins:403-407
                        s.start_idx += e.max_tries; s.end_idx += e.max_tries;
                        if (s.end_idx > len) s.end_idx = len;
                    }
                    s.s_idx = s.start_idx;
                    if (s.start_idx >= s.end_idx) { set_error(s, R3D_ERR_INDEX); s.phase = PH_ERROR; break; }
                    next_try = true;
                }
                if (next_try) {
                    next_try = false;
                    const int sidx = s.s_idx;
                    if (sidx >= len) { set_error(s, R3D_ERR_INDEX); s.phase = PH_ERROR; break; }     // od/ins:410
                    const int idx = list_entry(e, b, s.event - 1, ci, sidx, len);
                    if (idx < 0 || idx >= len) { set_error(s, R3D_ERR_INDEX); s.phase = PH_ERROR; break; }
                    const int obj = e.class_list[e.class_list_off[ci] + idx];
                    if (unpl[obj >> 5] & (1u << (obj & 31))) {       // od/ins:422-428
                        if (sidx == s.end_idx - 1) { s.remaining[ci] -= 1; new_slot = true; continue; }
                        s.s_idx = sidx + 1; next_try = true; continue;
                    }
                    s.cur_obj = obj; tryact = 1; s.phase = PH_AFTER_TRY;
                    break;
                }
            }
            if (s.phase != PH_DONE && s.phase != PH_ERROR && !tryact) { set_error(s, R3D_ERR_INDEX); s.phase = PH_ERROR; }
        }
        s.try_active = tryact; s.need_project = project; s.apply_flag = apply;
        s.n_feasible = 0; s.found_rank = INT_MAX; s.accepted = 0; s.chosen_rot = 0;
        e.gate_update[b] = project; e.gate_try[b] = tryact; e.gate_apply[b] = apply;
        for (int i = 0; i < 4; ++i) e.tickets[(size_t)b * 4 + i] = 0u;
        int* ac = e.active_count + 2 * slot;
        if (s.phase != PH_DONE && s.phase != PH_ERROR) atomicAdd(&ac[0], 1);
        // the last CTA publishes the number of unfinished scans straight into mapped host memory: the host polls this
        // word instead of waiting for a D2H copy (which would queue behind another engine's bulk transfers)
        __threadfence();
        if (atomicAdd(&ac[1], 1) == n_scans - 1) {
            const unsigned left = (unsigned)atomicAdd(&ac[0], 0);
            *(volatile unsigned long long*)(e.host_word + slot) = ((unsigned long long)(e.round_ctl[1] + round) << 32) | left;
            __threadfence_system();
            int* nx = e.active_count + 2 * ((slot + 32) & 63);      // re-arm the counters half a ring ahead
            nx[0] = 0; nx[1] = 0;
            e.work_cnt[0] = 0; e.work_cnt[1] = 0;                   // this round's work lists (filled by k_update)
            e.round_ctl[0] = round + 1u;                            // every other CTA has read it (it took its ticket)
        }
        if (tryact) atomicAdd(&e.stats[1], 1ull);
        if (apply) atomicAdd(&e.stats[2], 1ull);
    }
}

// ------------------------------------------------------------------------------- apply mask + min/max el
__device__ __forceinline__ bool pix_removed(const EngineDev& e, int b, const ScanState& s, int pix) {
    if (pix < 0) return false;
    if (e.dmask[(size_t)b * e.dwords + (pix >> 5)] & (1u << (pix & 31))) return true;
    // od/ins:486: an empty object pixel holds 500, so scene pixels farther than 500 count as covered
    return e.far_arr[b] && e.smooth[(size_t)b * e.hw + pix] > kEmptyRange;
}

// Slot update, one CTA per scan (A11/A12 + the decision how the range image is refreshed):
//  1. scene = scene[pix_id not in vis_px] (od/ins:488-501, 545).  vis_px lies inside the pixel rectangle select_emit
//     recorded, so only the points whose azimuth bin falls in that column range are visited (CSR by column, built once
//     per scan) plus the inserted tail: O(window) instead of O(N).  Notes whether a removed point held the scene's
//     min / max elevation.
//  2. full re-projection or in-place patch?  The image geometry (od/ins:97-98) depends only on the scene's min / max
//     elevation; if neither moved, every surviving point keeps its pixel and only the pixels of vis_px change.
//  3. patch: the z-buffer changes only at the pixels of vis_px — all their scene points were removed (od/ins:491)
//     and the visible object points were appended there (od/ins:545).
constexpr int UPDATE_THREADS = 256;
#ifndef R3D_UPDATE_G
#define R3D_UPDATE_G 8
#endif
constexpr int UPDATE_G = R3D_UPDATE_G;   // CTAs per scan; the last one to finish takes the decision and patches
__device__ __forceinline__ bool last_block_done(unsigned* ticket, unsigned n_blocks) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == n_blocks - 1;
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

__global__ void __launch_bounds__(UPDATE_THREADS) k_update(EngineDev e, int n_scans) {
    const int b = blockIdx.y;
    if (b >= n_scans) return;
    const int do_apply = e.gate_apply[b], do_update = e.gate_update[b];
    if (!do_apply && !do_update) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            e.gate_full[b] = 0; e.gate_patch[b] = 0;
            int* r = e.cf_rect + (size_t)b * 4; r[0] = 0; r[1] = -1; r[2] = 0; r[3] = -1;
        }
        return;
    }
    ScanState& s = e.st[b];
    const size_t base = (size_t)b * e.P;
    const int tid = blockIdx.x * UPDATE_THREADS + threadIdx.x, nthr = UPDATE_G * UPDATE_THREADS;
    bool extreme = false;
    if (do_apply) {
        const int* off = e.col_off + (size_t)b * (e.cols + 1);
        const int* idx = e.col_idx + (size_t)b * e.max_points;
        int c0 = s.d_c0, c1 = s.d_c1;
        if (e.far_arr[b]) { c0 = 0; c1 = e.cols - 1; }          // od/ins:486 quirk: covered pixels can be anywhere
        const unsigned long long lo = s.min_el_bits, hi = s.max_el_bits;
        if (c1 >= c0) {
            const int beg = c0 > 0 ? off[c0 - 1] : 0, end = off[c1];           // off[c] = END of column c's bucket
            for (int i = beg + tid; i < end; i += nthr) {
                const int p = idx[i];
                if (e.alive[base + p] && pix_removed(e, b, s, e.pix[base + p])) {
                    e.alive[base + p] = 0;
                    const unsigned long long bits = dbl_bits(e.el[base + p]);
                    extreme |= bits == lo || bits == hi;
                }
            }
        }
        for (int t = tid; t < s.tail_before; t += nthr) {
            const int p = s.n0 + t;
            if (e.alive[base + p] && pix_removed(e, b, s, e.pix[base + p])) {
                e.alive[base + p] = 0;
                const unsigned long long bits = dbl_bits(e.el[base + p]);
                extreme |= bits == lo || bits == hi;
            }
        }
    }
    if (__syncthreads_or(extreme) && threadIdx.x == 0) atomicOr(&s.extreme_removed, 1);
    if (!last_block_done(&e.tickets[(size_t)b * 4 + 0], UPDATE_G)) return;
    __shared__ int s_patch;
    if (threadIdx.x == 0) {
        int full = 0, patch = 0;
        int* rect = e.cf_rect + (size_t)b * 4;
        rect[0] = 0; rect[1] = -1; rect[2] = 0; rect[3] = -1;
        if (do_update) {
            const bool extended = s.new_min_bits < s.min_el_bits || s.new_max_bits > s.max_el_bits;
            full = s.first || *(volatile int*)&s.extreme_removed || extended || e.far_arr[b] || e.force_full;
            patch = !full;
            if (full) {
                s.min_el_bits = R3D_EMPTY_U64; s.max_el_bits = 0ull;
                rect[0] = 0; rect[1] = e.rows - 1; rect[2] = 0; rect[3] = e.cols - 1;
            } else {
                rect[0] = max(s.d_r0 - 4, 0); rect[1] = min(s.d_r1 + 4, e.rows - 1);
                rect[2] = max(s.d_c0 - 2, 0); rect[3] = min(s.d_c1 + 2, e.cols - 1);
            }
            s.first = 0;
            s.new_min_bits = R3D_EMPTY_U64; s.new_max_bits = 0ull;
            atomicAdd(&e.stats[full ? 0 : 3], 1ull);
        }
        s.extreme_removed = 0;
        e.gate_full[b] = full; e.gate_patch[b] = patch;
        s_patch = patch;
        if (full) e.full_list[atomicAdd(&e.work_cnt[0], 1)] = b;
        if (rect[1] >= rect[0] && rect[3] >= rect[2]) {             // close/fill tiles that overlap the rectangle
            const int ty0 = rect[0] / CF_TH, ty1 = rect[1] / CF_TH, tx0 = rect[2] / CF_TW, tx1 = rect[3] / CF_TW;
            const int nt = (ty1 - ty0 + 1) * (tx1 - tx0 + 1);
            int* task = e.cf_tasks + atomicAdd(&e.work_cnt[1], nt);
            for (int ty = ty0; ty <= ty1; ++ty)
                for (int tx = tx0; tx <= tx1; ++tx) *task++ = b * e.cf_tiles + ty * e.cf_tiles_x + tx;
        }
    }
    __syncthreads();
    if (do_update && e.task == 1) {
        const int ww = e.map_window * e.map_window / 32;
        unsigned* o = e.occ_win + (size_t)b * ww;
        for (int i = threadIdx.x; i < ww; i += UPDATE_THREADS) o[i] = 0u;
    }
    if (!s_patch) return;
    unsigned long long* z = e.zraw + (size_t)b * e.hw;
    const unsigned* dm = e.dmask + (size_t)b * e.dwords;
    if (s.d_r1 >= s.d_r0 && s.d_c1 >= s.d_c0) {
        // vis_px lies inside the rectangle select_emit recorded: visit only the mask words that hold its columns, row
        // by row (a word may be visited for two rows when the width is no multiple of 32; clearing twice is harmless)
        const int nw = (s.d_c1 >> 5) - (s.d_c0 >> 5) + 2, nrow = s.d_r1 - s.d_r0 + 1;
        for (int i = threadIdx.x; i < nrow * nw; i += UPDATE_THREADS) {
            const int r = s.d_r0 + i / nw;
            const int w = ((r * e.cols + s.d_c0) >> 5) + i % nw;
            if (w > ((r * e.cols + s.d_c1) >> 5)) continue;
            unsigned m = dm[w];
            while (m) { const int bit = __ffs(m) - 1; m &= m - 1; z[(w << 5) + bit] = R3D_EMPTY_U64; }
        }
    }
    __syncthreads();
    if (s.apply_flag)                                        // points appended by the accept being applied
        for (int p = s.n0 + s.tail_before + threadIdx.x; p < s.n0 + s.n_tail; p += UPDATE_THREADS)
            if (e.alive[base + p]) atomicMin(&z[e.pix[base + p]], dbl_bits(e.r[base + p]));
}

// A2 (od/ins:79-80) on the cached elevations: min / max over the live points (full path only)
__global__ void __launch_bounds__(STREAM_THREADS) k_minmax(EngineDev e, int n_scans) {
    __shared__ unsigned long long s_min[STREAM_THREADS / 32], s_max[STREAM_THREADS / 32];
    const int n_full = e.work_cnt[0];
    for (int li = blockIdx.y; li < n_full; li += gridDim.y) {       // the scans k_update listed for a full re-projection
    const int b = e.full_list[li];
    ScanState& s = e.st[b];
    const int n = s.n0 + s.n_tail;
    const int p0 = blockIdx.x * CHUNK;
    if (p0 >= n) continue;
    const size_t base = (size_t)b * e.P;
    unsigned long long lmin = R3D_EMPTY_U64, lmax = 0ull;
    for (int p = p0 + threadIdx.x; p < min(p0 + CHUNK, n); p += STREAM_THREADS) {
        if (e.alive[base + p]) {
            const unsigned long long bits = dbl_bits(e.el[base + p]);
            lmin = min(lmin, bits); lmax = max(lmax, bits);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lmin; s_max[threadIdx.x >> 5] = lmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < STREAM_THREADS / 32; ++w) { lmin = min(lmin, s_min[w]); lmax = max(lmax, s_max[w]); }
        if (lmax >= lmin) { atomicMin(&s.min_el_bits, lmin); atomicMax(&s.max_el_bits, lmax); }
    }
    __syncthreads();
    }
}

// clear the z-buffer of the scans that re-project in full and fix their image geometry
__global__ void __launch_bounds__(STREAM_THREADS) k_clear_images(EngineDev e, int n_scans) {
    const int n_full = e.work_cnt[0];
    for (int li = blockIdx.y; li < n_full; li += gridDim.y) {
        const int b = e.full_list[li];
        ScanState& s = e.st[b];
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            e.far_arr[b] = 0;
            if (s.min_el_bits == R3D_EMPTY_U64) { set_error(s, R3D_ERR_ASSERT); }
            s.geom = make_geom(e.rows, e.cols, e.cols, bits_dbl(s.max_el_bits), bits_dbl(s.min_el_bits));
        }
        unsigned long long* z = e.zraw + (size_t)b * e.hw;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.hw; i += gridDim.x * blockDim.x) z[i] = R3D_EMPTY_U64;
    }
}

// A3 (od/ins:85-130): bin every live point with the reference's truncation rule, write pix_id, 64-bit atomicMin of
// the range bits into the z-buffer.  Algorithmic traffic 20 B/point (+ 8 B/pixel for the z-buffer).
__global__ void __launch_bounds__(STREAM_THREADS) k_project(EngineDev e, int n_scans) {
    const int n_full = e.work_cnt[0];
    for (int li = blockIdx.y; li < n_full; li += gridDim.y) {
    const int b = e.full_list[li];
    ScanState& s = e.st[b];
    const int n = s.n0 + s.n_tail;
    const int p0 = blockIdx.x * CHUNK;
    if (p0 >= n) continue;
    const ImageGeom g = make_geom(e.rows, e.cols, e.cols, bits_dbl(s.max_el_bits), bits_dbl(s.min_el_bits));
    const size_t base = (size_t)b * e.P;           // P is a multiple of 16: the 4-point vectors below are aligned
    unsigned long long* z = e.zraw + (size_t)b * e.hw;
    auto one = [&](bool alive, double el, unsigned col, double r) -> int {
        if (!alive) return -1;
        const int row = bin_row(g, el);
        if (row < 0 || row >= g.rows) { set_error(s, R3D_ERR_ASSERT); return -1; }          // od/ins:111
        const int pix = row * g.cols + (int)col;
        atomicMin(&z[pix], dbl_bits(r));
        return pix;
    };
#pragma unroll 2
    for (int it = 0; it < CHUNK / (STREAM_THREADS * 4); ++it) {
        const int p = p0 + it * STREAM_THREADS * 4 + threadIdx.x * 4;
        if (p >= n) break;
        if (p + 3 < n) {                           // four points per thread: 4 + 32 + 8 + 32 bytes in, 16 out
            const uchar4 a = *reinterpret_cast<const uchar4*>(e.alive + base + p);
            if (!(a.x | a.y | a.z | a.w)) continue;
            const double2 e01 = *reinterpret_cast<const double2*>(e.el + base + p);
            const double2 e23 = *reinterpret_cast<const double2*>(e.el + base + p + 2);
            const ushort4 c = *reinterpret_cast<const ushort4*>(e.col + base + p);
            const double2 r01 = *reinterpret_cast<const double2*>(e.r + base + p);
            const double2 r23 = *reinterpret_cast<const double2*>(e.r + base + p + 2);
            int4 px;
            px.x = one(a.x, e01.x, c.x, r01.x); px.y = one(a.y, e01.y, c.y, r01.y);
            px.z = one(a.z, e23.x, c.z, r23.x); px.w = one(a.w, e23.y, c.w, r23.y);
            *reinterpret_cast<int4*>(e.pix + base + p) = px;
        } else {
            for (int q = p; q < n; ++q) e.pix[base + q] = one(e.alive[base + q], e.el[base + q], e.col[base + q], e.r[base + q]);
        }
    }
    }
}

struct RawImage {        // the engine's z-buffer as close/fill input
    const unsigned long long* raw;
    __device__ void load(int64_t i, double& v, uint8_t& o) const {
        const unsigned long long b = raw[i];
        const bool hit = b != R3D_EMPTY_U64;
        v = hit ? bits_dbl(b) : kEmptyRange;            // od/ins:100: empty = 500
        o = hit ? 3 : 0;
    }
    __device__ double lab(int64_t i) const { return raw[i] != R3D_EMPTY_U64 ? 1.0 : -1.0; }
};

// semseg addjust_map_2 (ss/ins:202-224): map cells (value != 0) that hold a live scene point with z < 1.5 and a
// non-ground label count as value 4 for this slot.  Kept as a per-scan bit window instead of rewriting the map.
__global__ void __launch_bounds__(STREAM_THREADS) k_adjust_map(EngineDev e, int n_scans) {
    const int b = blockIdx.y;
    if (b >= n_scans || !e.gate_update[b]) return;
    ScanState& s = e.st[b];
    const int n = s.n0 + s.n_tail;
    const int p0 = blockIdx.x * CHUNK;
    if (p0 >= n) return;
    const double* T = e.poses + (size_t)b * 16;
    const size_t base = (size_t)b * e.P;
    unsigned* o = e.occ_win + (size_t)b * (e.map_window * e.map_window / 32);
    for (int p = p0 + threadIdx.x; p < min(p0 + CHUNK, n); p += STREAM_THREADS) {
        if (!e.alive[base + p]) continue;
        const unsigned lab = e.label[base + p];
        bool ground = false;
        for (int i = 0; i < e.n_road_indexes; ++i) ground |= lab == (unsigned)e.road_indexes[i];
        if (ground) continue;
        double x, y, z;
        load_xyz(e, b, p, s.n0, x, y, z);
        if (!(z < 1.5)) continue;
        const double wx = add(add(add(mul(T[0], x), mul(T[1], y)), mul(T[2], z)), T[3]);
        const double wy = add(add(add(mul(T[4], x), mul(T[5], y)), mul(T[6], z)), T[7]);
        const int ix = trunc_to_int(sub(wx, (double)e.ss_move_x));
        const int iy = trunc_to_int(sub(wy, (double)e.ss_move_y));
        if (ix < 0 || iy < 0 || ix >= e.ss_sx || iy >= e.ss_sy) continue;   // reference: IndexError / wrap-around
        if (e.ss_map[(size_t)ix * e.ss_sy + iy] == 0) continue;
        const int lx = ix - s.win_x0, ly = iy - s.win_y0;
        if (lx < 0 || ly < 0 || lx >= e.map_window || ly >= e.map_window) { set_error(s, R3D_ERR_CAPACITY); continue; }
        const int bit = lx * e.map_window + ly;
        atomicOr(&o[bit >> 5], 1u << (bit & 31));
    }
}

// ------------------------------------------------------------------------------------------- placement
__device__ __forceinline__ bool surface_label(const ClassCfg& cc, unsigned lab) {
    bool ok = false;
    for (int i = 0; i < cc.n_surface; ++i) ok |= lab == (unsigned)cc.surface[i];
    return ok;
}

__device__ __forceinline__ int radius_index(const double* r2, double d2) {
    if (!(d2 <= r2[R3D_NUM_RADII - 1])) return R3D_NUM_RADII;
    int lo = 0, hi = R3D_NUM_RADII - 1;                   // smallest j with d2 <= r2[j]
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (d2 <= r2[mid]) hi = mid; else lo = mid + 1; }
    return lo;
}

// ------------------------------------------------------------------------------- road-level search grid
// The ORIGINAL scan never changes (correct_height reads original_pcl, od/fs:282), so its surface points (any label a
// class may stand on, z > -3) are bucketed ONCE per scan into a uniform grid (CSR by cell, rows contiguous in x).
// Points outside the grid extent are clamped into border cells: distances are always computed from the coordinates,
// and clamping never increases a cell-index difference, so the square searches below stay exact.
__device__ __forceinline__ bool any_surface_label(const EngineDev& e, unsigned lab) {
    for (int c = 0; c < e.n_classes; ++c) {
        const ClassCfg& cc = e.classes[c];
        for (int i = 0; i < cc.n_surface; ++i) if (lab == (unsigned)cc.surface[i]) return true;
    }
    return false;
}
__device__ __forceinline__ int grid_coord(const EngineDev& e, float v) {
    const int i = (int)floorf(v * e.grid_inv_cell) + (e.G >> 1);
    return max(0, min(i, e.G - 1));
}

template <int PASS>     // 1: count per cell, 2: scatter (after the prefix scan)
__global__ void __launch_bounds__(STREAM_THREADS) k_grid_build(EngineDev e, int n_scans) {
    const int b = blockIdx.y;
    if (b >= n_scans) return;
    const int n0 = e.st[b].n0;
    const int p0 = blockIdx.x * CHUNK;
    if (p0 >= n0) return;
    const size_t base = (size_t)b * e.P;
    int* cell = e.gcell + (size_t)b * e.G * e.G;
    int* acell = e.acell + (size_t)b * e.G * e.G;
    float4* out = e.gpts + (size_t)b * e.max_points;
    float4* aout = e.apts + (size_t)b * e.max_points;
    for (int p = p0 + threadIdx.x; p < min(p0 + CHUNK, n0); p += STREAM_THREADS) {
        const float4 v = __ldg(&e.xyzi[(size_t)b * e.max_points + p]);
        const unsigned lab = e.label[base + p];
        const int c = grid_coord(e, v.y) * e.G + grid_coord(e, v.x);
        if (!(e.task == 0 && lab == (unsigned)e.road_label)) {      // OD: Road points are never obstacles (od/ins:353-355)
            if (PASS == 1) atomicAdd(&acell[c], 1);
            else aout[atomicAdd(&acell[c], 1)] = make_float4(v.x, v.y, v.z, __int_as_float(p));
        }
        if (!((double)v.z > -3.0) || !any_surface_label(e, lab)) continue;       // od/fs:154-155
        if (PASS == 1) atomicAdd(&cell[c], 1);
        else out[atomicAdd(&cell[c], 1)] = make_float4(v.x, v.y, v.z, __uint_as_float(lab));
    }
}

// Chebyshev distance (in cells, capped at NEAR_CAP) from every cell to the nearest cell that holds a surface point:
// the road-level search of a candidate whose surroundings are empty starts at that ring instead of growing through
// the empty ones.  Separable: row pass (min |dx| along the row), then column pass (min over dy of max(|dy|, row value)).
constexpr int NEAR_CAP = 12;
template <int PASS>
__global__ void __launch_bounds__(256) k_grid_near(EngineDev e, int n_scans) {
    const int b = blockIdx.y, G = e.G;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (b >= n_scans || c >= G * G) return;
    const int y = c / G, x = c % G;
    const size_t gb = (size_t)b * G * G;
    int best = NEAR_CAP;
    if (PASS == 1) {
        const int* cell = e.gcell + gb;
        for (int dx = -(NEAR_CAP - 1); dx <= NEAR_CAP - 1; ++dx) {
            const int x1 = x + dx;
            if (x1 < 0 || x1 >= G) continue;
            const int q = y * G + x1;
            if (cell[q] > (q > 0 ? cell[q - 1] : 0)) best = min(best, abs(dx));
        }
        e.gscratch[gb + c] = (unsigned char)best;
    } else {
        for (int dy = -(NEAR_CAP - 1); dy <= NEAR_CAP - 1; ++dy) {
            const int y1 = y + dy;
            if (y1 < 0 || y1 >= G) continue;
            best = min(best, max(abs(dy), (int)e.gscratch[gb + (size_t)y1 * G + x]));
        }
        e.gnear[gb + c] = (unsigned char)best;
    }
}

// One more once-per-scan CSR index over the ORIGINAL points (their azimuth bin never changes): by image column,
// for k_apply_window.
template <int PASS>     // 1: count, 2: scatter point indices
__global__ void __launch_bounds__(STREAM_THREADS) k_index_build(EngineDev e, int n_scans) {
    const int b = blockIdx.y;
    if (b >= n_scans) return;
    const int n0 = e.st[b].n0;
    const int p0 = blockIdx.x * CHUNK;
    if (p0 >= n0) return;
    const size_t base = (size_t)b * e.P;
    int* coff = e.col_off + (size_t)b * (e.cols + 1);
    int* cidx = e.col_idx + (size_t)b * e.max_points;
    for (int p = p0 + threadIdx.x; p < min(p0 + CHUNK, n0); p += STREAM_THREADS) {
        const int c = e.col[base + p];
        if (PASS == 1) atomicAdd(&coff[c], 1);
        else cidx[atomicAdd(&coff[c], 1)] = p;
    }
}

// exclusive prefix sum of the per-cell counts (one CTA per scan, four cells per thread and step); after the scatter
// pass cell[c] = END of cell c
__global__ void __launch_bounds__(1024) k_bucket_scan(int* arr, size_t stride, int n, int n_scans) {
    const int b = blockIdx.x;
    if (b >= n_scans) return;
    int* cell = arr + (size_t)b * stride;
    __shared__ int s_w[32];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool vec = (stride & 3) == 0;                  // rows of `arr` 16-byte aligned: int4 loads / stores
    for (int i0 = 0; i0 < n; i0 += 4096) {
        const int i = i0 + threadIdx.x * 4;
        int v[4] = {0, 0, 0, 0};
        if (vec && i + 3 < n) { const int4 t = *reinterpret_cast<const int4*>(cell + i); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else for (int u = 0; u < 4; ++u) if (i + u < n) v[u] = cell[i + u];
        const int mine = v[0] + v[1] + v[2] + v[3];
        int inc = mine;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        int off = s_run, tot = 0;
        for (int j = 0; j < 32; ++j) { if (j < w) off += s_w[j]; tot += s_w[j]; }
        int e0 = off + inc - mine;                        // exclusive prefix of this thread's first cell
        int o4[4];
        for (int u = 0; u < 4; ++u) { o4[u] = e0; e0 += v[u]; }
        if (vec && i + 3 < n) *reinterpret_cast<int4*>(cell + i) = make_int4(o4[0], o4[1], o4[2], o4[3]);
        else for (int u = 0; u < 4; ++u) if (i + u < n) cell[i + u] = o4[u];
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
}

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

__device__ __forceinline__ unsigned group_mask() { return 0xFFu << ((threadIdx.x & 31) & ~(GRP - 1)); }

// Visit every point stored in the grid cells [x0, x1] x [y0, y1] MINUS the cells of the hole [hx0, hx1] x [hy0, hy1]
// (a rectangle inside the first one; hx1 < hx0 = no hole) with the 8 lanes of a group.  The lanes fetch the CSR
// ranges of up to 8 rows at once (a row of cells is contiguous; a row crossing the hole has a left and a right
// segment), then stride over each segment with four independent 16-byte loads in flight per lane: the walk is a
// chain of L2 latencies, so memory-level parallelism is what counts.  `f(v)` is called per point; `stop()` is polled
// after every row (group-uniform early exit).
struct CellRect { int x0, x1, y0, y1; };
template <class F, class S>
__device__ __forceinline__ void group_visit(const int* __restrict__ cell, const float4* __restrict__ pts, int G, CellRect rc,
                                            CellRect hole, int gl, unsigned gm, F f, S stop) {
    for (int yb = rc.y0; yb <= rc.y1; yb += GRP) {
        const int nrows = min(GRP, rc.y1 - yb + 1);
        int beg_a = 0, end_a = 0, beg_b = 0, end_b = 0;
        if (gl < nrows) {
            const int y = yb + gl, row = y * G;
            const bool split = hole.x1 >= hole.x0 && y >= hole.y0 && y <= hole.y1;
            const int xa1 = split ? hole.x0 - 1 : rc.x1;                 // segment A: [x0, xa1], B: [hole.x1 + 1, x1]
            if (xa1 >= rc.x0) {
                const int c0 = row + rc.x0;
                beg_a = c0 > 0 ? __ldg(&cell[c0 - 1]) : 0;
                end_a = __ldg(&cell[row + xa1]);
            }
            if (split && hole.x1 < rc.x1) {
                beg_b = __ldg(&cell[row + hole.x1]);
                end_b = __ldg(&cell[row + rc.x1]);
            }
        }
        for (int r = 0; r < nrows; ++r) {
#pragma unroll
            for (int seg = 0; seg < 2; ++seg) {
                const int rb = __shfl_sync(gm, seg ? beg_b : beg_a, r, GRP), re = __shfl_sync(gm, seg ? end_b : end_a, r, GRP);
                for (int p = rb + gl; p < re; p += 4 * GRP) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int q = p + u * GRP; if (q < re) v[u] = __ldg(&pts[q]); }
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (p + u * GRP < re) f(v[u]);
                }
            }
            if (stop()) return;
        }
    }
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
__device__ bool group_road_level(const EngineDev& e, int b, const SurfaceSet& surf, double cx, double cy, int gl, unsigned gm,
                                 double& level) {
    const int G = e.G;
    const int* __restrict__ cell = e.gcell + (size_t)b * G * G;
    const float4* __restrict__ pts = e.gpts + (size_t)b * e.max_points;
    double best = 1e300;
    CellRect hole{0, -1, 0, -1};
    const double step = e.grid_cell;
    double R = 0.5 * step;
    // every cell closer than `ring` cells (Chebyshev) to the centre's cell is empty: start at the radius whose square
    // of cells still lies inside that empty block and treat the block as already visited
    const int ring = e.gnear[(size_t)b * G * G + (size_t)grid_coord(e, (float)cy) * G + grid_coord(e, (float)cx)];
    if (ring >= 2) {
        R = fmin((ring - 1) * step, 5.0);
        hole = cells_within(e, cx, cy, R - 0.01);
    }
    for (;; R = fmin(R < step ? step : R + step, 5.0)) {
        const CellRect rc = cells_within(e, cx, cy, R);
        group_visit(cell, pts, G, rc, hole, gl, gm, [&](const float4& v) {
            if (!in_surface(surf, __float_as_uint(v.w))) return;
            const double dx = sub((double)v.x, cx), dy = sub((double)v.y, cy);
            best = fmin(best, add(mul(dx, dx), mul(dy, dy)));                           // od/fs:153
        }, [] { return false; });
        for (int o = GRP / 2; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(gm, best, o));
        if (best <= R * R || R >= 5.0) break;          // every point outside the scanned cells is farther than R
        hole = rc;
    }
    if (!(best <= e.radii_sq[R3D_NUM_RADII - 1])) return false;
    const int j = radius_index(e.radii_sq, best);       // smallest j with best <= r_j^2
    if (!e.radii_ok[j]) return false;                   // od/fs:156-160: no surface within reach
    const double r2 = e.radii_sq[j];
    long long zsum = 0;
    int cnt = 0;
    group_visit(cell, pts, G, cells_within(e, cx, cy, sqrt(r2)), CellRect{0, -1, 0, -1}, gl, gm, [&](const float4& v) {
        if (!in_surface(surf, __float_as_uint(v.w))) return;
        const double dx = sub((double)v.x, cx), dy = sub((double)v.y, cy);
        if (add(mul(dx, dx), mul(dy, dy)) <= r2) { zsum += __double2ll_rn(mul((double)v.z, kFix)); ++cnt; }
    }, [] { return false; });
    for (int o = GRP / 2; o > 0; o >>= 1) {
        zsum += __shfl_xor_sync(gm, zsum, o);
        cnt += __shfl_xor_sync(gm, cnt, o);
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

// rare paths of the collision test, kept out of line so the common path stays small (per-lane results):
// obstacle points among the points of one already inserted object (its tail slice)
__device__ __noinline__ bool tail_hits_candidate(const EngineDev& e, int b, const ScanState& s, const ClassCfg& cc,
                                                 const YawTest& yt, double zmin_ped, int t0, int cnt, int gl) {
    const size_t base = (size_t)b * e.P;
    const bool ped = cc.pedestrian != 0;
    for (int i = gl; i < cnt; i += GRP) {
        const size_t t = (size_t)b * e.max_inserted + t0 + i;
        const double x = e.tail_x[t], y = e.tail_y[t], z = e.tail_z[t];
        if ((!ped || z >= zmin_ped) && inside_yaw(yt, x, y, z) && obstacle_point(e, b, s, cc, base, s.n0 + t0 + i))
            return true;
    }
    return false;
}
// any object point of the candidate strictly inside a scene box (od/fs:129-134)
__device__ __noinline__ bool object_in_scene_box(const BoxTest* box_test, const double* ox, const double* oy, const double* oz,
                                                 int count, double c, double sn, double dz, int gl) {
    const BoxTest sbt = *box_test;
    for (int i = gl; i < count; i += 2 * GRP) {
        const int i1 = i + GRP;
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
__device__ bool group_collides(const EngineDev& e, int b, const ScanState& s, const ObjBox& ob, const ClassCfg& cc, double c,
                               double sn, double level, int gl, unsigned gm) {
    const YawBox yb = make_yaw_box(ob.cx, ob.cy, ob.a, ob.b, c, sn);
    const YawTest yt = make_yaw_test(yb, level, ob.length, ob.width, ob.height);
    const double zmin_ped = add(level, 0.1);                                      // od/fs:123-124
    const bool ped = cc.pedestrian != 0;
    const size_t base = (size_t)b * e.P;
    const int gshift = (threadIdx.x & 31) & ~(GRP - 1);
    bool hit = false;
    {
        const int G = e.G;
        const float fcx = (float)yb.cx, fcy = (float)yb.cy, fr = (float)ob.reach + 1e-3f, fr2 = fr * fr;
        const float zlo = (float)level - 1e-3f, zhi = (float)(level + ob.height) + 1e-3f;
        const CellRect rc{grid_coord(e, fcx - fr), grid_coord(e, fcx + fr), grid_coord(e, fcy - fr), grid_coord(e, fcy + fr)};
        group_visit(e.acell + (size_t)b * G * G, e.apts + (size_t)b * e.max_points, G, rc, CellRect{0, -1, 0, -1}, gl, gm,
                    [&](const float4& v) {
            if (hit) return;
            const float dx = v.x - fcx, dy = v.y - fcy;                  // cheap conservative pruning first
            if (dx * dx + dy * dy > fr2 || v.z < zlo || v.z > zhi) return;
            const double x = v.x, yy = v.y, z = v.z;
            if (ped && !(z >= zmin_ped)) return;
            if (!inside_yaw(yt, x, yy, z)) return;                       // exact test (cb:30-66)
            if (obstacle_point(e, b, s, cc, base, __float_as_int(v.w))) hit = true;
        }, [&] { return (__ballot_sync(gm, hit) & gm) != 0u; });
        if (__ballot_sync(gm, hit) & gm) return true;
    }
    // the lanes look at 8 boxes at a time; only the boxes whose bounding circle reaches the candidate's are tested
    const int nbox0 = s.n_boxes - s.n_inserted;
    int t_run = 0;
    for (int j0 = 0; j0 < s.n_inserted; j0 += GRP) {                     // the tail of placed object j lies inside its box
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
        for (int o = 1; o < GRP; o <<= 1) { const int t = __shfl_up_sync(gm, inc, o, GRP); if (gl >= o) inc += t; }
        unsigned m = (__ballot_sync(gm, near) & gm) >> gshift;
        while (m) {
            const int q = __ffs(m) - 1; m &= m - 1;
            const int qcnt = __shfl_sync(gm, cnt, q, GRP), qt0 = t_run + __shfl_sync(gm, inc, q, GRP) - qcnt;
            hit = tail_hits_candidate(e, b, s, cc, yt, zmin_ped, qt0, qcnt, gl);
            if (__ballot_sync(gm, hit) & gm) return true;
        }
        t_run += __shfl_sync(gm, inc, GRP - 1, GRP);
    }
    const double dz = sub(level, ob.cz);
    const double *ox = e.obj_x + ob.first, *oy = e.obj_y + ob.first, *oz = e.obj_z + ob.first;
    for (int b0 = 0; b0 < s.n_boxes; b0 += GRP) {                        // (ii) od/fs:129-134
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
            hit = object_in_scene_box(&e.box_tests[(size_t)b * e.max_boxes + b0 + q], ox, oy, oz, ob.count, c, sn, dz, gl);
            if (__ballot_sync(gm, hit) & gm) return true;
        }
    }
    return false;
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

// stage 2b (semseg): A6b (ss/fs:231-248) with the reference's carried z shift (ss/fs:146-147 is in place): the yaws
// are visited in order by one CTA; world = T . [x y z 1] - move, astype(int); every in-map cell value must be allowed.
__global__ void __launch_bounds__(1024) k_onmap_ss(EngineDev e, int n_scans) {
    const int b = blockIdx.x;
    if (b >= n_scans || !e.gate_try[b]) return;
    const ScanState& s = e.st[b];
    const ObjBox ob = e.try_obj[b];
    const unsigned okmask = e.classes[ob.cls].map_ok_mask;
    const double* T = e.poses + (size_t)b * 16;
    const double t00 = T[0], t01 = T[1], t02 = T[2], t03 = T[3], t10 = T[4], t11 = T[5], t12 = T[6], t13 = T[7];
    const unsigned* o = e.occ_win + (size_t)b * (e.map_window * e.map_window / 32);
    const size_t cb = (size_t)b * (e.K + 1);
    double dz = 0.0;
    for (int k = 1; k <= e.K; ++k) {
        const double c = e.cos_k[k], sn = e.sin_k[k];
        int bad = 0;
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x) {
            const double x0 = e.obj_x[ob.first + i], y0 = e.obj_y[ob.first + i];
            const double x = sub(mul(c, x0), mul(sn, y0)), y = add(mul(sn, x0), mul(c, y0));
            const double z = add(e.obj_z[ob.first + i], dz);
            const double wx = add(add(add(mul(t00, x), mul(t01, y)), mul(t02, z)), t03);
            const double wy = add(add(add(mul(t10, x), mul(t11, y)), mul(t12, z)), t13);
            const int ix = trunc_to_int(sub(wx, (double)e.ss_move_x));
            const int iy = trunc_to_int(sub(wy, (double)e.ss_move_y));
            if (ix < e.ss_sx && ix > -1 && iy < e.ss_sy && iy > -1) {
                unsigned v = e.ss_map[(size_t)ix * e.ss_sy + iy];
                const int lx = ix - s.win_x0, ly = iy - s.win_y0;
                if (lx >= 0 && ly >= 0 && lx < e.map_window && ly < e.map_window) {
                    const int bit = lx * e.map_window + ly;
                    if (o[bit >> 5] & (1u << (bit & 31))) v = 4;
                }
                if (!((okmask >> v) & 1u)) bad = 1;
            }
        }
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

// ------------------------------------------------------------------------------------------- occlusion
// object point i of candidate k -> (pix, r) in the CURRENT scene geometry (od/ins:474-478, sample=True)
struct ObjProj { double x, y, z, r, el; int pix, col; };
__device__ __forceinline__ ObjProj project_obj_point(const EngineDev& e, const ObjBox& ob, const ImageGeom& g, int i,
                                                     double c, double sn, double dz, ScanState& s) {
    ObjProj o;
    const double x0 = e.obj_x[ob.first + i], y0 = e.obj_y[ob.first + i];
    o.x = sub(mul(c, x0), mul(sn, y0));
    o.y = add(mul(sn, x0), mul(c, y0));
    o.z = add(e.obj_z[ob.first + i], dz);
    o.r = range3(o.x, o.y, o.z);
    o.el = elevation(o.z, o.r);
    const int row = bin_row(g, o.el);
    o.col = bin_col(g, azimuth(o.x, o.y));
    o.pix = -1;
    if (row >= 0 && row < g.rows) {                                    // od/ins:108-109
        if (o.col < 0 || o.col >= g.cols) set_error(s, R3D_ERR_ASSERT);  // od/ins:113
        else o.pix = row * g.cols + o.col;
    }
    return o;
}

// A11 (od/ins:486-501) for every feasible candidate: V = number of object points whose pixel is visible.  A pixel
// that holds object points keeps its own min range through smooth_out, and min_r < scene <=> some point of the pixel
// has r < scene, so no z-buffer is needed for the count: pass 1 marks visible pixels in a shared-memory bit image,
// pass 2 counts the points on marked pixels.  Candidates are visited in rotation order with an ordered early-out
// (the reference stops at the first candidate that keeps >= min_points, od/ins:530-561).
__global__ void k_phase_gate(EngineDev e, int n_scans) {           // which scans found nothing in the first window
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_scans) e.need2[b] = e.gate_try[b] && e.st[b].found_rank == INT_MAX && e.n_list[b] > e.cand_window;
}

__global__ void __launch_bounds__(128) k_occl_count(EngineDev e, int n_scans, int phase) {
    const int b = blockIdx.y;
    if (b >= n_scans || !(phase == 2 ? e.need2[b] : e.gate_try[b])) return;
    ScanState& s = e.st[b];
    extern __shared__ unsigned s_bits[];
    __shared__ int s_cnt, s_stop;
    __shared__ int s_warp[4];
    // ordered list of the feasible rotations (the order find_possible_places returns them, od/fs:288-296): every CTA
    // of the scan compacts the flag bytes itself; the first one publishes the list for k_select_emit
    unsigned short* s_feas = reinterpret_cast<unsigned short*>(s_bits + e.dwords);
    const size_t cb = (size_t)b * (e.K + 1);
    const int nf = block_compact(e.cand_flags + cb, e.K, CF_ONMAP | CF_HOK | CF_COLLIDE, CF_ONMAP | CF_HOK, s_feas, s_warp);
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < nf; i += blockDim.x) e.feas[(size_t)b * e.K + i] = s_feas[i];
        if (threadIdx.x == 0) s.n_feasible = nf;
    }
    if ((int)blockIdx.x >= nf) return;
    for (int i = threadIdx.x; i < e.dwords; i += blockDim.x) s_bits[i] = 0u;
    const ObjBox ob = e.try_obj[b];
    const int min_pts = e.classes[ob.cls].min_points;
    const ImageGeom g = s.geom;
    const double* smooth = e.smooth + (size_t)b * e.hw;
    int* pixbuf = e.occ_pix + ((size_t)b * OCC_G + blockIdx.x) * e.max_obj_points;
    __syncthreads();
    for (int rank = blockIdx.x; rank < nf; rank += gridDim.x) {
        if (threadIdx.x == 0) { s_stop = *(volatile int*)&s.found_rank < rank; s_cnt = 0; }
        __syncthreads();
        if (s_stop) break;                                        // an earlier candidate already passed
        const int k = s_feas[rank];
        const double c = e.cos_k[k], sn = e.sin_k[k];
        const double dz = sub(e.cand_level[cb + k], ob.cz);
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x) {
            const ObjProj o = project_obj_point(e, ob, g, i, c, sn, dz, s);
            pixbuf[i] = o.pix;
            if (o.pix >= 0 && o.r < smooth[o.pix]) atomicOr(&s_bits[o.pix >> 5], 1u << (o.pix & 31));
        }
        __syncthreads();
        int cnt = 0;
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x) {
            const int pix = pixbuf[i];
            if (pix >= 0 && (s_bits[pix >> 5] & (1u << (pix & 31)))) ++cnt;
        }
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x) {
            const int pix = pixbuf[i];
            if (pix >= 0) s_bits[pix >> 5] = 0u;
        }
        if (threadIdx.x == 0) {
            e.cand_v[cb + k] = s_cnt;
            if (s_cnt > 0 && s_cnt >= min_pts) atomicMin(&s.found_rank, rank);      // od/ins:530-536
        }
        __syncthreads();
    }
}

// smoothed object range at pixel (r, c): own min range, or the neighbour mean where the 5x3 closing switches an
// empty pixel on (cl:26-62).  `dil` is the bit image of the dilated occupancy (union of the 5x3 neighbourhoods of the
// object's pixels), so closed(q) = AND of dil over the in-image 5x3 neighbourhood of q.
__device__ bool obj_pixel_value(const unsigned long long* raw, const unsigned* dil, int H, int W, int r, int c, double& val) {
    const unsigned long long own = raw[r * W + c];
    if (own != R3D_EMPTY_U64) { val = bits_dbl(own); return true; }
    for (int dr = -2; dr <= 2; ++dr)
        for (int dc = -1; dc <= 1; ++dc) {
            const int r1 = r + dr, c1 = c + dc;
            if (r1 < 0 || r1 >= H || c1 < 0 || c1 >= W) continue;          // outside the image: ignored by the erosion
            const int q = r1 * W + c1;
            if (!(dil[q >> 5] & (1u << (q & 31)))) return false;
        }
    int neighbors = 0;
    double sum = 0.0;
    for (int dr = -2; dr <= 2; ++dr)
        for (int dc = -1; dc <= 1; ++dc) {
            const int r1 = r + dr, c1 = c + dc;
            if (r1 < 0 || r1 >= H || c1 < 0 || c1 >= W) continue;
            const unsigned long long v = raw[r1 * W + c1];
            if (v != R3D_EMPTY_U64) { neighbors += 1; sum = add(sum, bits_dbl(v)); }
        }
    if (neighbors == 0) return false;
    val = __ddiv_rn(sum, (double)neighbors);
    return true;
}

__device__ void bitonic_sort_u64(unsigned long long* keys, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], bb = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > bb) == up) { keys[i] = bb; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

// A11 + A12 for the chosen candidate of each scan: the first feasible rotation that keeps >= min_points (accepted),
// else the last feasible one (its vis_px still deletes scene points in the reference, od/ins:472-501).  Builds the
// candidate's z-buffer in a scratch image, closes / fills it around the object, compares with the scene image
// (strict <) into the vis_px bit mask, and on acceptance appends the visible object points in (pix_id, index) order
// to the scene tail, the `check` record and the scene boxes.
#ifndef R3D_SEL_TILE_PX
#define R3D_SEL_TILE_PX 8192
#endif
constexpr int SEL_TILE_PX = R3D_SEL_TILE_PX;      // pixels of the shared-memory object tile (8 B of fp64 range each)

// local variant of obj_pixel_value on the shared-memory tile (rows r_lo.., cols c_lo.., nr x nc).  Pixels outside the
// tile but inside the image hold no object point and are farther than the 5x3 window from every object pixel, so
// their occupancy and dilation are 0; pixels outside the image are ignored by the erosion.
__device__ __forceinline__ bool tile_pixel_value(const unsigned long long* tile, const unsigned* dil, int H, int W, int r_lo,
                                                 int c_lo, int nr, int nc, int r, int c, double& val) {
    const unsigned long long own = tile[(r - r_lo) * nc + (c - c_lo)];
    if (own != R3D_EMPTY_U64) { val = bits_dbl(own); return true; }
    for (int dr = -2; dr <= 2; ++dr)
        for (int dc = -1; dc <= 1; ++dc) {
            const int r1 = r + dr, c1 = c + dc;
            if (r1 < 0 || r1 >= H || c1 < 0 || c1 >= W) continue;
            const int lr = r1 - r_lo, lc = c1 - c_lo;
            if (lr < 0 || lr >= nr || lc < 0 || lc >= nc) return false;
            const int q = lr * nc + lc;
            if (!(dil[q >> 5] & (1u << (q & 31)))) return false;
        }
    int neighbors = 0;
    double sum = 0.0;
    for (int dr = -2; dr <= 2; ++dr)
        for (int dc = -1; dc <= 1; ++dc) {
            const int lr = r + dr - r_lo, lc = c + dc - c_lo;
            if (lr < 0 || lr >= nr || lc < 0 || lc >= nc) continue;
            const unsigned long long v = tile[lr * nc + lc];
            if (v != R3D_EMPTY_U64) { neighbors += 1; sum = add(sum, bits_dbl(v)); }
        }
    if (neighbors == 0) return false;
    val = __ddiv_rn(sum, (double)neighbors);
    return true;
}

// A11 + A12 for the chosen candidate of each scan: the first feasible rotation that keeps >= min_points (accepted),
// else the last feasible one (its vis_px still deletes scene points in the reference, od/ins:472-501).  Builds the
// candidate's z-buffer (in a shared-memory tile around the object when it fits, else in a global scratch image),
// closes / fills it, compares with the scene image (strict <) into the vis_px bit mask, and on acceptance appends the
// visible object points in (pix_id, index) order to the scene tail, the `check` record and the scene boxes.
__global__ void __launch_bounds__(512) k_select_emit(EngineDev e, int n_scans, int key_cap, int smem_pts) {
    const int b = blockIdx.x;
    if (b >= n_scans || !e.gate_try[b]) return;
    ScanState& s = e.st[b];
    const int nf = s.n_feasible;
    if (nf == 0) return;
    extern __shared__ unsigned long long s_dyn[];
    // sort keys / ranges / pixel ids of the object's points: shared memory for objects up to smem_pts points, the
    // per-scan global scratch for larger ones (trucks with > 10k points)
    unsigned long long* s_keys = s_dyn;                                   // [key_cap]
    double* s_r = reinterpret_cast<double*>(s_dyn + key_cap);            // [smem_pts]
    unsigned long long* s_tile = s_dyn + key_cap + smem_pts;             // [SEL_TILE_PX]
    int* s_pix = reinterpret_cast<int*>(s_tile + SEL_TILE_PX);           // [smem_pts]
    unsigned* s_dil = reinterpret_cast<unsigned*>(s_pix + smem_pts);     // [SEL_TILE_PX / 32]
    unsigned* s_vis = s_dil + SEL_TILE_PX / 32;                          // [SEL_TILE_PX / 32]
    if (e.try_obj[b].count > smem_pts) {
        s_keys = e.sel_keys + (size_t)b * e.sel_key_cap;
        s_r = e.sel_r + (size_t)b * e.max_obj_points;
        s_pix = e.sel_pix + (size_t)b * e.max_obj_points;
    }
    __shared__ int s_nvis;
    __shared__ int s_rect[4];
    __shared__ unsigned long long s_el[2];
    if (threadIdx.x == 0) {
        s_rect[0] = INT_MAX; s_rect[1] = -1; s_rect[2] = INT_MAX; s_rect[3] = -1; s_el[0] = R3D_EMPTY_U64; s_el[1] = 0ull;
        s_nvis = 0;
    }
    const bool accepted = s.found_rank < nf;
    const int rank = accepted ? s.found_rank : nf - 1;
    const int k = e.feas[(size_t)b * e.K + rank];
    const ObjBox ob = e.obj[s.cur_obj];
    const ImageGeom g = s.geom;
    const int H = g.rows, W = g.cols;
    const size_t cb = (size_t)b * (e.K + 1);
    const double c = e.cos_k[k], sn = e.sin_k[k];
    const double level = e.cand_level[cb + k];
    const double dz = sub(level, ob.cz);
    unsigned* dm = e.dmask + (size_t)b * e.dwords;
    const double* smooth = e.smooth + (size_t)b * e.hw;
    const int t0 = s.n_tail, chk0 = s.n_check, nbox0 = s.n_boxes, nins0 = s.n_inserted, n0 = s.n0;
    for (int i = threadIdx.x; i < e.dwords; i += blockDim.x) dm[i] = 0u;
    __syncthreads();
    // project every object point once (od/ins:474-478); pixel rectangle of the object
    {
        int r_lo = INT_MAX, r_hi = -1, c_lo = INT_MAX, c_hi = -1;
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x) {
            const ObjProj o = project_obj_point(e, ob, g, i, c, sn, dz, s);
            s_pix[i] = o.pix; s_r[i] = o.r;
            if (o.pix < 0) continue;
            const int pr = o.pix / W, pc = o.pix % W;
            r_lo = min(r_lo, pr); r_hi = max(r_hi, pr); c_lo = min(c_lo, pc); c_hi = max(c_hi, pc);
        }
        for (int o = 16; o > 0; o >>= 1) {
            r_lo = min(r_lo, __shfl_xor_sync(0xffffffffu, r_lo, o)); r_hi = max(r_hi, __shfl_xor_sync(0xffffffffu, r_hi, o));
            c_lo = min(c_lo, __shfl_xor_sync(0xffffffffu, c_lo, o)); c_hi = max(c_hi, __shfl_xor_sync(0xffffffffu, c_hi, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&s_rect[0], r_lo); atomicMax(&s_rect[1], r_hi); atomicMin(&s_rect[2], c_lo); atomicMax(&s_rect[3], c_hi);
        }
    }
    __syncthreads();
    const bool any_px = s_rect[1] >= 0;
    // vis_px can only lie within the object's pixels grown by the 5x3 window
    const int wr0 = any_px ? max(s_rect[0] - 2, 0) : 0, wr1 = any_px ? min(s_rect[1] + 2, H - 1) : -1;
    const int wc0 = any_px ? max(s_rect[2] - 1, 0) : 0, wc1 = any_px ? min(s_rect[3] + 1, W - 1) : -1;
    const int nr = wr1 - wr0 + 1, nc = wc1 - wc0 + 1;
    const bool in_smem = any_px && nr * nc <= SEL_TILE_PX;
    if (threadIdx.x == 0 && any_px) atomicAdd(&e.stats[in_smem ? 4 : 5], 1ull);
    if (in_smem) {
        const int npx = nr * nc;
        for (int i = threadIdx.x; i < npx; i += blockDim.x) s_tile[i] = R3D_EMPTY_U64;
        for (int i = threadIdx.x; i < (npx + 31) / 32; i += blockDim.x) { s_dil[i] = 0u; s_vis[i] = 0u; }
        __syncthreads();
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x) {
            const int pix = s_pix[i];
            if (pix < 0) continue;
            const int pr = pix / W, pc = pix % W;
            atomicMin(&s_tile[(pr - wr0) * nc + (pc - wc0)], dbl_bits(s_r[i]));
            for (int dr = -2; dr <= 2; ++dr)                             // dilated occupancy (5 rows x 3 cols)
                for (int dc = -1; dc <= 1; ++dc) {
                    const int r1 = pr + dr, c1 = pc + dc;
                    if (r1 < 0 || r1 >= H || c1 < 0 || c1 >= W) continue;
                    const int q = (r1 - wr0) * nc + (c1 - wc0);
                    atomicOr(&s_dil[q >> 5], 1u << (q & 31));
                }
        }
        __syncthreads();
        // Only a pixel inside the dilated occupancy can survive the closing, and the dilation bits cover exactly the
        // 5x3 neighbourhoods of the object's pixels: one visit per such pixel of the tile (instead of one per
        // (object point, neighbour) pair, which evaluated most pixels many times).
        for (int lq = threadIdx.x; lq < npx; lq += blockDim.x) {
            if (!(s_dil[lq >> 5] & (1u << (lq & 31)))) continue;
            const int r = wr0 + lq / nc, cc = wc0 + lq % nc;
            double val;
            if (tile_pixel_value(s_tile, s_dil, H, W, wr0, wc0, nr, nc, r, cc, val) && val < smooth[r * W + cc]) {   // od/ins:486
                atomicOr(&s_vis[lq >> 5], 1u << (lq & 31));
                const int q = r * W + cc;
                atomicOr(&dm[q >> 5], 1u << (q & 31));
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x) {
            const int pix = s_pix[i];
            if (pix < 0) continue;
            const int lq = (pix / W - wr0) * nc + (pix % W - wc0);
            if (s_vis[lq >> 5] & (1u << (lq & 31))) {
                const int slot = atomicAdd(&s_nvis, 1);
                s_keys[slot] = ((unsigned long long)(unsigned)pix << 32) | (unsigned)i;
            }
        }
    } else if (any_px) {
        // object too wide for the tile (very close / very large): global scratch image + global dilation mask
        unsigned long long* raw = e.obj_raw + (size_t)b * e.hw;
        unsigned* vm = e.vmask + (size_t)b * e.dwords;
        for (int i = threadIdx.x; i < e.dwords; i += blockDim.x) vm[i] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x) {
            const int pix = s_pix[i];
            if (pix < 0) continue;
            atomicMin(&raw[pix], dbl_bits(s_r[i]));
            const int pr = pix / W, pc = pix % W;
            for (int dr = -2; dr <= 2; ++dr)
                for (int dc = -1; dc <= 1; ++dc) {
                    const int r1 = pr + dr, c1 = pc + dc;
                    if (r1 < 0 || r1 >= H || c1 < 0 || c1 >= W) continue;
                    const int q = r1 * W + c1;
                    atomicOr(&vm[q >> 5], 1u << (q & 31));
                }
        }
        __threadfence_block();
        __syncthreads();
        for (int t = threadIdx.x; t < ob.count * 15; t += blockDim.x) {
            const int i = t / 15, o = t % 15;
            const int pix = s_pix[i];
            if (pix < 0) continue;
            const int r = pix / W + (o / 3 - 2), cc = pix % W + (o % 3 - 1);
            if (r < 0 || r >= H || cc < 0 || cc >= W) continue;
            const int q = r * W + cc;
            double val;
            if (obj_pixel_value(raw, vm, H, W, r, cc, val) && val < smooth[q]) atomicOr(&dm[q >> 5], 1u << (q & 31));   // od/ins:486
        }
        __threadfence_block();
        __syncthreads();
        for (int i = threadIdx.x; i < ob.count; i += blockDim.x) {
            const int pix = s_pix[i];
            if (pix >= 0 && (dm[pix >> 5] & (1u << (pix & 31)))) {
                const int slot = atomicAdd(&s_nvis, 1);
                s_keys[slot] = ((unsigned long long)(unsigned)pix << 32) | (unsigned)i;
            }
            if (pix >= 0) raw[pix] = R3D_EMPTY_U64;                      // leave the scratch z-buffer empty
        }
    }
    __syncthreads();
    // visible object points, ordered by (pix_id, original index) as the reference's per-pixel loop emits them
    const int nvis = s_nvis;
    if (accepted) {
        int np2 = 1;
        while (np2 < nvis) np2 <<= 1;
        for (int i = nvis + threadIdx.x; i < np2; i += blockDim.x) s_keys[i] = R3D_EMPTY_U64;
        __syncthreads();
        bitonic_sort_u64(s_keys, np2);
        if (t0 + nvis > e.max_inserted || nbox0 + 1 > e.max_boxes || nins0 + 1 > e.max_events) {
            __syncthreads();
            if (threadIdx.x == 0) { set_error(s, R3D_ERR_CAPACITY); s.phase = PH_ERROR; }
        } else {
            const size_t base = (size_t)b * e.P + n0 + t0;
            const size_t tb = (size_t)b * e.max_inserted + t0;
            const size_t chk = ((size_t)b * e.max_inserted + chk0) * 5;
            for (int j = threadIdx.x; j < nvis; j += blockDim.x) {
                const int i = (int)(s_keys[j] & 0xffffffffull);
                const ObjProj o = project_obj_point(e, ob, g, i, c, sn, dz, s);
                e.tail_x[tb + j] = o.x; e.tail_y[tb + j] = o.y; e.tail_z[tb + j] = o.z;
                const float inten = e.obj_i[ob.first + i];
                const unsigned lab = e.obj_label[ob.first + i];
                e.tail_i[tb + j] = inten;
                e.label[base + j] = lab;
                e.r[base + j] = o.r; e.el[base + j] = o.el;
                atomicMin(&s_el[0], dbl_bits(o.el)); atomicMax(&s_el[1], dbl_bits(o.el));
                e.col[base + j] = (unsigned short)o.col;
                e.pix[base + j] = o.pix;
                e.alive[base + j] = 1;
                float* ck = e.check + chk + (size_t)j * 5;
                ck[0] = (float)o.x; ck[1] = (float)o.y; ck[2] = (float)o.z; ck[3] = inten; ck[4] = (float)lab;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int* rec = e.inserted + ((size_t)b * e.max_events + nins0) * 4;
                rec[0] = s.cur_obj; rec[1] = k; rec[2] = ob.cls; rec[3] = nvis;
                const YawBox yb = make_yaw_box(ob.cx, ob.cy, ob.a, ob.b, c, sn);
                double* ib = e.inserted_box + ((size_t)b * e.max_events + nins0) * 8;
                ib[0] = yb.cx; ib[1] = yb.cy; ib[2] = level; ib[3] = yb.m00; ib[4] = yb.m10;
                ib[5] = ob.length; ib[6] = ob.width; ib[7] = ob.height;
                Box bx = yaw_box_to_box(yb, level, ob.length, ob.width, ob.height);      // od/ins:555
                bx.reach = ob.reach;
                e.boxes[(size_t)b * e.max_boxes + nbox0] = bx;
                e.box_tests[(size_t)b * e.max_boxes + nbox0] = make_box_test(bx);
                s.n_boxes = nbox0 + 1; s.n_inserted = nins0 + 1;
                s.tail_before = t0; s.n_tail = t0 + nvis; s.n_check = chk0 + nvis;
                s.new_min_bits = s_el[0]; s.new_max_bits = s_el[1];
            }
        }
    }
    if (threadIdx.x == 0) {
        s.accepted = accepted ? 1 : 0; s.chosen_rot = k; s.chosen_v = nvis;
        s.d_r0 = wr0; s.d_r1 = wr1; s.d_c0 = wc0; s.d_c1 = wc1;
        if (!accepted) { s.new_min_bits = R3D_EMPTY_U64; s.new_max_bits = 0ull; }
    }
}

// --------------------------------------------------------------------------------------------- outputs
// A14 (od/ds:76-109, ss/ds:72-106): surviving rows in order (original points, then inserted points), cast to
// float32 / uint32.  40 B/point of algorithmic traffic.
// Three launches: per-chunk live counts (grid = chunks x scans), offsets (one CTA: totals per scan, exclusive scan over
// the scans, chunk offsets), chunk-wise stable compaction.
__global__ void __launch_bounds__(STREAM_THREADS) k_out_count(EngineDev e, int n_scans) {
    const int b = blockIdx.y;
    if (b >= n_scans) return;
    const ScanState& s = e.st[b];
    const int n = s.n0 + s.n_tail;
    const int p0 = blockIdx.x * CHUNK;
    const size_t base = (size_t)b * e.P;
    int cnt = 0;
    if (p0 < n) {
        const int p1 = min(p0 + CHUNK, n);
        if (p1 - p0 == CHUNK && ((base + p0) & 15) == 0) {     // whole aligned chunk: 16 alive bytes per thread in one load
            const uint4 v = *reinterpret_cast<const uint4*>(e.alive + base + p0 + threadIdx.x * 16);
            cnt = __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) + __popc(v.w & 0x01010101u);
        } else {
            for (int p = p0 + threadIdx.x; p < p1; p += STREAM_THREADS) cnt += e.alive[base + p] ? 1 : 0;
        }
    }
    __shared__ int s_w[STREAM_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < STREAM_THREADS / 32; ++w) t += s_w[w];
        e.chunk_cnt[(size_t)b * e.max_chunks + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_out_offsets(EngineDev e, int n_scans, int chunks) {
    __shared__ long long s_w[32], s_c[32];
    __shared__ long long s_run, s_crun;
    if (threadIdx.x == 0) { s_run = 0; s_crun = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < n_scans; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        long long tot = 0, chk = 0;
        if (b < n_scans) {
            const int* cc = e.chunk_cnt + (size_t)b * e.max_chunks;
            for (int c = 0; c < chunks; ++c) tot += cc[c];
            chk = e.st[b].n_check;
        }
        long long inc = tot, cinc = chk;
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, inc, o), u = __shfl_up_sync(0xffffffffu, cinc, o);
            if (lane >= o) { inc += t; cinc += u; }
        }
        if (lane == 31) { s_w[w] = inc; s_c[w] = cinc; }
        __syncthreads();
        long long off = s_run, coff = s_crun;
        for (int i = 0; i < w; ++i) { off += s_w[i]; coff += s_c[i]; }
        if (b < n_scans) {
            e.out_off[b] = off + inc - tot; e.check_off[b] = coff + cinc - chk; e.out_count[b] = tot;
            if (b == n_scans - 1) { e.out_off[n_scans] = off + inc; e.check_off[n_scans] = coff + cinc; }
        }
        __syncthreads();
        if (threadIdx.x == 1023) { s_run = off + inc; s_crun = coff + cinc; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(STREAM_THREADS) k_out_write(EngineDev e, int n_scans) {
    const int b = blockIdx.y;
    if (b >= n_scans) return;
    const ScanState& s = e.st[b];
    const int n = s.n0 + s.n_tail;
    const int p0 = blockIdx.x * CHUNK;
    const size_t base = (size_t)b * e.P;
    if (blockIdx.x == 0) {                              // the `check` record of the scan (od/ds:91-93)
        const long long c0 = e.check_off[b];
        const float* ck = e.check + (size_t)b * e.max_inserted * 5;
        for (int i = threadIdx.x; i < s.n_check * 5; i += blockDim.x) e.out_check[c0 * 5 + i] = ck[i];
    }
    if (p0 >= n) return;
    long long o0 = e.out_off[b];
    {
        const int* cc = e.chunk_cnt + (size_t)b * e.max_chunks;
        for (int c = 0; c < (int)blockIdx.x; ++c) o0 += cc[c];
    }
    __shared__ int s_w[STREAM_THREADS / 32];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int q0 = p0; q0 < min(p0 + CHUNK, n); q0 += STREAM_THREADS) {
        const int p = q0 + threadIdx.x;
        const bool a = p < n && e.alive[base + p];
        float4 v;
        unsigned lab = 0;
        if (a) {                                        // issue the loads before the scan's barriers
            if (p < s.n0) v = __ldg(&e.xyzi[(size_t)b * e.max_points + p]);
            else {
                const size_t t = (size_t)b * e.max_inserted + (p - s.n0);
                v = make_float4((float)e.tail_x[t], (float)e.tail_y[t], (float)e.tail_z[t], e.tail_i[t]);
            }
            lab = e.label[base + p];
        }
        const unsigned m = __ballot_sync(0xffffffffu, a);
        if (lane == 0) s_w[w] = __popc(m);
        __syncthreads();
        int off = s_run, tot = 0;
        for (int i = 0; i < STREAM_THREADS / 32; ++i) { if (i < w) off += s_w[i]; tot += s_w[i]; }
        if (a) {
            const long long o = o0 + off + __popc(m & ((1u << lane) - 1u));
            e.out_xyzi[o] = v;
            e.out_label[o] = lab;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
    }
}

}  // namespace r3d
