// Rich-map generation of the object-detection pipeline:
// object_detection/rich_map/single_drivable_area_map.py:123-194 (abbreviated od/rm), batched over frames.
// Per frame: 1 m grid over the scan's xy extent; road points rasterised; closing(disk(4)) -> road map; the
// 8-neighbour ring around the road, dilation(disk(2)) -> pedestrian-area map.  One CTA per frame walks the passes
// with the (tiny, ~160 x 160) maps ping-ponging through global scratch that stays in L1 / L2.
#include <algorithm>
#include "r3d_common.cuh"
#include "r3d_host.h"
#include "../../include/real3d_b200.h"

namespace {

constexpr int RM_THREADS = 1024;

// od/rm:123-133: min / max of x and y over ALL points; int() truncates toward zero
__global__ void __launch_bounds__(RM_THREADS) k_rm_extents(const float4* __restrict__ xyzi, const long long* __restrict__ pt_off,
                                                           int n_scans, int* __restrict__ dims) {
    const int b = blockIdx.x;
    if (b >= n_scans) return;
    const long long o = pt_off[b];
    const int n = (int)(pt_off[b + 1] - o);
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = threadIdx.x; i < n; i += RM_THREADS) {
        const float4 v = __ldg(&xyzi[o + i]);
        mnx = fminf(mnx, v.x); mxx = fmaxf(mxx, v.x); mny = fminf(mny, v.y); mxy = fmaxf(mxy, v.y);
    }
    for (int s = 16; s > 0; s >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, s)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, s));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, s)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, s));
    }
    __shared__ float s_v[4][RM_THREADS / 32];
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        s_v[0][w] = mnx; s_v[1][w] = mxx; s_v[2][w] = mny; s_v[3][w] = mxy;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RM_THREADS / 32; ++w) {
            mnx = fminf(mnx, s_v[0][w]); mxx = fmaxf(mxx, s_v[1][w]); mny = fminf(mny, s_v[2][w]); mxy = fmaxf(mxy, s_v[3][w]);
        }
        int* d = dims + (size_t)b * 4;
        if (n == 0) { d[0] = d[1] = d[2] = d[3] = 0; return; }
        const int min_x = (int)mnx, min_y = (int)mny;                  // od/rm:123-124
        const int max_x = (int)mxx + 1, max_y = (int)mxy + 1;          // od/rm:126-127
        d[0] = max_x - min_x; d[1] = max_y - min_y; d[2] = min_x; d[3] = min_y;
    }
}

// binary max (dilation) / min (erosion) over a disk footprint, cells outside the image ignored
template <int R, bool DILATE>
__device__ __forceinline__ void disk_pass(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int sx, int sy) {
    for (int c = threadIdx.x; c < sx * sy; c += RM_THREADS) {
        const int r0 = c / sy, c0 = c % sy;
        bool v = !DILATE;
        for (int dr = -R; dr <= R; ++dr) {
            const int r1 = r0 + dr;
            if (r1 < 0 || r1 >= sx) continue;
            for (int dc = -R; dc <= R; ++dc) {
                if (dr * dr + dc * dc > R * R) continue;               // skimage.morphology.disk
                const int c1 = c0 + dc;
                if (c1 < 0 || c1 >= sy) continue;
                const bool a = in[r1 * sy + c1] != 0;
                if (DILATE) v |= a; else v &= a;
            }
        }
        out[c] = v ? 1 : 0;
    }
}

__global__ void __launch_bounds__(RM_THREADS) k_rm_build(const float4* __restrict__ xyzi, const unsigned* __restrict__ labels,
                                                         const long long* __restrict__ pt_off, int n_scans, unsigned road_label,
                                                         const int* __restrict__ dims, const long long* __restrict__ map_off,
                                                         unsigned char* __restrict__ road_out, unsigned char* __restrict__ ped_out,
                                                         unsigned char* __restrict__ scratch, long long total_cells) {
    const int b = blockIdx.x;
    if (b >= n_scans) return;
    const int sx = dims[b * 4 + 0], sy = dims[b * 4 + 1], min_x = dims[b * 4 + 2], min_y = dims[b * 4 + 3];
    const int cells = sx * sy;
    if (cells <= 0) return;
    const long long mo = map_off[b];
    unsigned char* A = scratch + mo;
    unsigned char* B = scratch + total_cells + mo;
    unsigned char* road = road_out + mo;
    unsigned char* ped = ped_out + mo;
    for (int c = threadIdx.x; c < cells; c += RM_THREADS) A[c] = 0;
    __syncthreads();
    const long long o = pt_off[b];
    const int n = (int)(pt_off[b + 1] - o);
    for (int i = threadIdx.x; i < n; i += RM_THREADS) {                 // od/rm:136-145
        if (labels[o + i] != road_label) continue;
        const float4 v = __ldg(&xyzi[o + i]);
        const int ix = (int)((double)v.x - (double)min_x), iy = (int)((double)v.y - (double)min_y);
        if (ix >= 0 && ix < sx && iy >= 0 && iy < sy) A[ix * sy + iy] = 1;
    }
    __syncthreads();
    disk_pass<4, true>(A, B, sx, sy);                                   // od/rm:150-156 closing(disk(4)) = erode(dilate)
    __syncthreads();
    disk_pass<4, false>(B, road, sx, sy);
    __syncthreads();
    for (int c = threadIdx.x; c < cells; c += RM_THREADS) {             // od/rm:164-180: not road, some 8-neighbour is road
        const int r0 = c / sy, c0 = c % sy;
        bool near_road = false;
        if (!road[c])
            for (int dr = -1; dr <= 1; ++dr)
                for (int dc = -1; dc <= 1; ++dc) {
                    const int r1 = r0 + dr, c1 = c0 + dc;
                    if (r1 >= 0 && r1 < sx && c1 >= 0 && c1 < sy && road[r1 * sy + c1]) near_road = true;
                }
        A[c] = near_road ? 1 : 0;
    }
    __syncthreads();
    disk_pass<2, true>(A, ped, sx, sy);                                 // od/rm:182-188 dilation(disk(2))
}


// ------------------------------------------------------------------------------------------- semantic segmentation
// semantic_segmentation/rich_map/drivable_area_map.py:122-206 (abbreviated ss/rm): ONE map per sequence.  Every frame is
// moved to the world frame (points = T . [x y z 1], ss/rm:134-137), the 1 m grid spans the xy extent of all frames
// (floor of the minimum, int() + 1 of the maximum, ss/rm:158-166) and the surface points are rasterised IN ORDER
// (frames, then points): road labels write 1, parking labels 2 unless the cell already holds 3; sidewalk labels
// write 3 for good (ss/rm:190-200).  So a cell ends as 3 if any sidewalk point fell in it, else as the class of the
// LAST road / parking point that fell in it: an order-independent atomicMax over keys
// (global point order + 1) << 2 | class, with all-ones for the sticky 3.
constexpr int RMS_THREADS = 256;
constexpr int RMS_CHUNK = 4096;          // points per CTA
constexpr int RMS_MAX_LABELS = 32;

// monotone u64 image of a double (total order of the finite values)
__device__ __forceinline__ unsigned long long ord_bits(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct Pose34 { double t[16]; };
__device__ __forceinline__ Pose34 load_pose(const double* __restrict__ poses, int f) {
    Pose34 p;
#pragma unroll
    for (int i = 0; i < 16; ++i) p.t[i] = poses[(size_t)f * 16 + i];
    return p;
}
// row r of T . [x y z 1] as numpy's `t_matrix @ points.T` (ss/rm:136) evaluates it: the BLAS dgemm micro-kernel keeps
// one accumulator per output element and walks k = 0..3 with fused multiply-adds (acc = fma(t_k, p_k, acc) from 0).
// Verified against numpy 2.3 / OpenBLAS 0.3.30 in the build container: 6000 of 6000 sampled coordinates bit-equal to
// this chain (5675 of 6000 to the unfused sum).
__device__ __forceinline__ double pose_row(const Pose34& p, int r, double x, double y, double z) {
    return __fma_rn(p.t[4 * r + 3], 1.0, __fma_rn(p.t[4 * r + 2], z, __fma_rn(p.t[4 * r + 1], y, __dmul_rn(p.t[4 * r], x))));
}

// ss/rm:130-146: world-frame xy extent of ALL points of the frames; ext[4] = ord_bits of {min x, max x, min y, max y}
__global__ void __launch_bounds__(RMS_THREADS) k_rms_extents(const float4* __restrict__ xyzi, const long long* __restrict__ pt_off,
                                                             const double* __restrict__ poses, int n_frames,
                                                             unsigned long long* __restrict__ ext) {
    const int f = blockIdx.y;
    if (f >= n_frames) return;
    const long long o = pt_off[f];
    const int n = (int)(pt_off[f + 1] - o);
    const int p0 = blockIdx.x * RMS_CHUNK;
    if (p0 >= n) return;
    const Pose34 T = load_pose(poses, f);
    double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = p0 + threadIdx.x; i < min(p0 + RMS_CHUNK, n); i += RMS_THREADS) {
        const float4 v = __ldg(&xyzi[o + i]);
        const double w = pose_row(T, 3, v.x, v.y, v.z);                                       // ss/rm:137
        const double wx = __ddiv_rn(pose_row(T, 0, v.x, v.y, v.z), w), wy = __ddiv_rn(pose_row(T, 1, v.x, v.y, v.z), w);
        mnx = fmin(mnx, wx); mxx = fmax(mxx, wx); mny = fmin(mny, wy); mxy = fmax(mxy, wy);
    }
    for (int s = 16; s > 0; s >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, s)); mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, s));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, s)); mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, s));
    }
    if ((threadIdx.x & 31) == 0 && mnx <= mxx) {
        atomicMin(&ext[0], ord_bits(mnx)); atomicMax(&ext[1], ord_bits(mxx));
        atomicMin(&ext[2], ord_bits(mny)); atomicMax(&ext[3], ord_bits(mxy));
    }
}

__global__ void k_rms_extents_init(unsigned long long* ext) {
    ext[0] = ~0ull; ext[1] = 0ull; ext[2] = ~0ull; ext[3] = 0ull;
}
// ss/rm:158-166: min = int(floor(min)), max = int(max) + 1; out = {min_x, min_y, size_x, size_y, any point}
__global__ void k_rms_extents_finish(const unsigned long long* ext, long long* out) {
    auto dec = [](unsigned long long o) { return __longlong_as_double((long long)((o >> 63) ? (o & 0x7fffffffffffffffull) : ~o)); };
    if (ext[0] == ~0ull) { out[0] = out[1] = out[2] = out[3] = out[4] = 0; return; }
    const long long min_x = (long long)floor(dec(ext[0])), min_y = (long long)floor(dec(ext[2]));
    const long long max_x = (long long)dec(ext[1]) + 1, max_y = (long long)dec(ext[3]) + 1;
    out[0] = min_x; out[1] = min_y; out[2] = max_x - min_x; out[3] = max_y - min_y; out[4] = 1;
}

// ss/rm:172-200 for the frames of one call; `order_base` = number of points of the frames rasterised before
__global__ void __launch_bounds__(RMS_THREADS) k_rms_raster(const float4* __restrict__ xyzi, const unsigned* __restrict__ labels,
                                                            const long long* __restrict__ pt_off, const double* __restrict__ poses,
                                                            int n_frames, const int* __restrict__ surf_label,
                                                            const int* __restrict__ surf_class, int n_surf, long long min_x,
                                                            long long min_y, int sx, int sy, long long order_base,
                                                            unsigned long long* __restrict__ keymap, int* __restrict__ err) {
    const int f = blockIdx.y;
    if (f >= n_frames) return;
    __shared__ unsigned s_lab[RMS_MAX_LABELS];
    __shared__ int s_cls[RMS_MAX_LABELS];
    if (threadIdx.x < RMS_MAX_LABELS) {
        s_lab[threadIdx.x] = threadIdx.x < n_surf ? (unsigned)surf_label[threadIdx.x] : 0xFFFFFFFFu;
        s_cls[threadIdx.x] = threadIdx.x < n_surf ? surf_class[threadIdx.x] : 0;
    }
    __syncthreads();
    const long long o = pt_off[f];
    const int n = (int)(pt_off[f + 1] - o);
    const int p0 = blockIdx.x * RMS_CHUNK;
    if (p0 >= n) return;
    const Pose34 T = load_pose(poses, f);
    const double dminx = (double)min_x, dminy = (double)min_y;
    const int p1 = min(p0 + RMS_CHUNK, n);
    const int lane = threadIdx.x & 31;
    for (int i0 = p0; i0 < p1; i0 += RMS_THREADS) {              // warp-uniform trip count (the warp votes below)
        const int i = i0 + threadIdx.x;
        int cls = 0;
        long long cell = -1;
        if (i < p1) {
            const unsigned lab = __ldg(&labels[o + i]);
            for (int j = 0; j < n_surf; ++j) if (lab == s_lab[j]) { cls = s_cls[j]; break; }   // ss/rm:185
        }
        if (cls) {
            const float4 v = __ldg(&xyzi[o + i]);
            const double w = pose_row(T, 3, v.x, v.y, v.z);                                       // ss/rm:178
            const double px = __dsub_rn(__ddiv_rn(pose_row(T, 0, v.x, v.y, v.z), w), dminx);
            const double py = __dsub_rn(__ddiv_rn(pose_row(T, 1, v.x, v.y, v.z), w), dminy);
            if (px < 0.0 || py < 0.0) atomicExch(err, 1);                                      // ss/rm:190 assert
            else {
                const long long ix = (long long)px, iy = (long long)py;                         // int(): truncation
                if (ix >= sx || iy >= sy) atomicExch(err, 2);                                  // IndexError in the reference
                else cell = ix * sy + iy;
            }
        }
        // Neighbouring points of a LiDAR ring fall into the same 1 m cell, and along a drive hundreds of frames hit
        // it: one atomic per (warp, cell) instead of one per point.  Within a warp the lane order is the point order,
        // so the group's largest key is its highest lane's — unless a sticky class-3 point is in the group.
        const unsigned active = __ballot_sync(0xffffffffu, cell >= 0);
        if (cell >= 0) {
            const unsigned group = __match_any_sync(active, (unsigned long long)cell);
            const unsigned sticky = __ballot_sync(group, cls == 3);
            if (lane == 31 - __clz(group)) {
                const unsigned long long key = sticky ? ~0ull
                    : ((unsigned long long)(order_base + (o - pt_off[0]) + i + 1) << 2) | (unsigned)cls;
                atomicMax(&keymap[cell], key);
            }
        }
    }
}

__global__ void k_rms_finalize(const unsigned long long* __restrict__ keymap, unsigned char* __restrict__ map, long long cells) {
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = keymap[c];
        map[c] = k == 0ull ? 0 : (k == ~0ull ? 3 : (unsigned char)(k & 3ull));
    }
}

}  // namespace

extern "C" int r3d_rich_map_ss_extents(const float* xyzi, const int64_t* point_offsets, const double* poses, int32_t n_frames,
                                       int32_t max_points, int32_t first_call, uint64_t* ext_state, int64_t* out5,
                                       r3d_stream stream_) {
    if (!xyzi || !point_offsets || !poses || !ext_state || n_frames <= 0 || max_points < 0)
        return r3d_fail(R3D_ERR_ARG, "r3d_rich_map_ss_extents: bad argument");
    cudaStream_t st = (cudaStream_t)stream_;
    if (first_call) { k_rms_extents_init<<<1, 1, 0, st>>>((unsigned long long*)ext_state); r3d_count_launch(); }
    const int chunks = (max_points + RMS_CHUNK - 1) / RMS_CHUNK;
    if (chunks > 0) {
        k_rms_extents<<<dim3(chunks, n_frames), RMS_THREADS, 0, st>>>((const float4*)xyzi, (const long long*)point_offsets, poses,
                                                                        n_frames, (unsigned long long*)ext_state);
        r3d_count_launch();
    }
    if (out5) { k_rms_extents_finish<<<1, 1, 0, st>>>((const unsigned long long*)ext_state, (long long*)out5); r3d_count_launch(); }
    return r3d_check_launch("r3d_rich_map_ss_extents");
}

extern "C" int r3d_rich_map_ss_raster(const float* xyzi, const uint32_t* labels, const int64_t* point_offsets, const double* poses,
                                      int32_t n_frames, int32_t max_points, const int32_t* surface_labels,
                                      const int32_t* surface_classes, int32_t n_surface, int64_t min_x, int64_t min_y,
                                      int32_t size_x, int32_t size_y, int64_t order_base, uint64_t* keymap, int32_t* error_flag,
                                      r3d_stream stream_) {
    if (!xyzi || !labels || !point_offsets || !poses || !surface_labels || !surface_classes || !keymap || !error_flag ||
        n_frames <= 0 || n_surface < 0 || n_surface > RMS_MAX_LABELS || size_x <= 0 || size_y <= 0 || max_points < 0)
        return r3d_fail(R3D_ERR_ARG, "r3d_rich_map_ss_raster: bad argument");
    cudaStream_t st = (cudaStream_t)stream_;
    const int chunks = (max_points + RMS_CHUNK - 1) / RMS_CHUNK;
    if (chunks > 0) {
        k_rms_raster<<<dim3(chunks, n_frames), RMS_THREADS, 0, st>>>((const float4*)xyzi, labels, (const long long*)point_offsets,
                                                                       poses, n_frames, surface_labels, surface_classes, n_surface,
                                                                       (long long)min_x, (long long)min_y, size_x, size_y,
                                                                       (long long)order_base, (unsigned long long*)keymap, error_flag);
        r3d_count_launch();
    }
    return r3d_check_launch("r3d_rich_map_ss_raster");
}

extern "C" int r3d_rich_map_ss_finalize(const uint64_t* keymap, int64_t cells, uint8_t* map_out, r3d_stream stream_) {
    if (!keymap || !map_out || cells <= 0) return r3d_fail(R3D_ERR_ARG, "r3d_rich_map_ss_finalize: bad argument");
    cudaStream_t st = (cudaStream_t)stream_;
    const int blocks = (int)std::min<long long>((cells + 255) / 256, 148 * 8);
    k_rms_finalize<<<blocks, 256, 0, st>>>((const unsigned long long*)keymap, map_out, (long long)cells);
    r3d_count_launch();
    return r3d_check_launch("r3d_rich_map_ss_finalize");
}

extern "C" int r3d_rich_map_od_extents(const float* xyzi, const int64_t* point_offsets, int32_t n_scans, int32_t* dims,
                                       r3d_stream stream_) {
    if (!xyzi || !point_offsets || !dims || n_scans <= 0) return r3d_fail(R3D_ERR_ARG, "r3d_rich_map_od_extents: bad argument");
    cudaStream_t st = (cudaStream_t)stream_;
    k_rm_extents<<<n_scans, RM_THREADS, 0, st>>>((const float4*)xyzi, (const long long*)point_offsets, n_scans, dims);
    r3d_count_launch();
    return r3d_check_launch("r3d_rich_map_od_extents");
}

extern "C" int r3d_rich_map_od_build(const float* xyzi, const uint32_t* labels, const int64_t* point_offsets, int32_t n_scans,
                                     uint32_t road_label, const int32_t* dims, const int64_t* map_offsets, int64_t total_cells,
                                     uint8_t* road_out, uint8_t* ped_out, uint8_t* scratch, r3d_stream stream_) {
    if (!xyzi || !labels || !point_offsets || !dims || !map_offsets || !road_out || !ped_out || !scratch || n_scans <= 0)
        return r3d_fail(R3D_ERR_ARG, "r3d_rich_map_od_build: bad argument");
    cudaStream_t st = (cudaStream_t)stream_;
    k_rm_build<<<n_scans, RM_THREADS, 0, st>>>((const float4*)xyzi, labels, (const long long*)point_offsets, n_scans, road_label,
                                               dims, (const long long*)map_offsets, road_out, ped_out, scratch, total_cells);
    r3d_count_launch();
    return r3d_check_launch("r3d_rich_map_od_build");
}
