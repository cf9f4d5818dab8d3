// Rich-map generation of the object-detection pipeline:
// object_detection/rich_map/single_drivable_area_map.py:123-194 (abbreviated od/rm), batched over frames.
// Per frame: 1 m grid over the scan's xy extent; road points rasterised; closing(disk(4)) -> road map; the
// 8-neighbour ring around the road, dilation(disk(2)) -> pedestrian-area map.  One CTA per frame walks the passes
// with the (tiny, ~160 x 160) maps ping-ponging through global scratch that stays in L1 / L2.
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

}  // namespace

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
