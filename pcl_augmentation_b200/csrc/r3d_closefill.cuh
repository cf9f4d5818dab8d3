// A4: class_closing + smooth_out (cl:9-62) as one tiled stencil kernel.
//   closing(rectangle(5, 3)) = 5 rows x 3 cols dilate then erode on the binary occupancy, pixels outside the image
//   ignored, no azimuth wrap; pixels switched on by the closing that hold no point get the mean of their occupied
//   neighbours' ORIGINAL ranges, summed in the reference's (drow, dcol) order in fp64 (cl:46-51,57).
// Templated on the input adaptor so the same code serves the float64 train/label pair of the drop-in smooth_out and
// the engine's raw uint64 z-buffer.  The tile (+ a 4-row / 2-col halo) is staged once in shared memory with
// coalesced row reads; dilation, erosion and the hole fill then run out of shared memory, so HBM sees one read of
// the input (x1.33 for the halo, mostly L2 hits) and one write of the output: 16 B/pixel for the engine variant.
#pragma once
#include "r3d_common.cuh"

constexpr int CF_TH = 32;       // output rows per CTA
constexpr int CF_TW = 64;       // output cols per CTA
constexpr int CF_THREADS = 256;
constexpr int CF_HR = 4, CF_HC = 2;                    // halo: two 5x3 passes
constexpr int CF_SH = CF_TH + 2 * CF_HR;               // staged rows
constexpr int CF_SW = CF_TW + 2 * CF_HC;               // staged cols

// one CF_TH x CF_TW output tile at (r0, c0) of image z; every thread of the CTA calls it (it synchronises)
template <class In>
__device__ __forceinline__ void close_fill_tile(const In& in, int H, int W, int64_t base, int r0, int c0, int z,
                                                double* __restrict__ out_train, double* __restrict__ out_label,
                                                uint8_t* __restrict__ closed_out, int* __restrict__ far_flag) {
    __shared__ double s_val[CF_SH][CF_SW];
    __shared__ uint8_t s_occ[CF_SH][CF_SW];            // bit0: occupancy (clip(label,0,1) > 0), bit1: label == 1
    __shared__ uint8_t s_dil[CF_TH + 4][CF_TW + 2];
    for (int i = threadIdx.x; i < CF_SH * CF_SW; i += CF_THREADS) {
        const int lr = i / CF_SW, lc = i % CF_SW;
        const int r = r0 - CF_HR + lr, c = c0 - CF_HC + lc;
        uint8_t o = 0;
        double v = 0.0;
        if (r >= 0 && r < H && c >= 0 && c < W) in.load(base + (int64_t)r * W + c, v, o);
        s_val[lr][lc] = v;
        s_occ[lr][lc] = o;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (CF_TH + 4) * (CF_TW + 2); i += CF_THREADS) {
        const int lr = i / (CF_TW + 2), lc = i % (CF_TW + 2);
        const int r = r0 - 2 + lr, c = c0 - 1 + lc;
        uint8_t d = 1;                                   // outside the image: neutral for the erosion
        if (r >= 0 && r < H && c >= 0 && c < W) {
            d = 0;
#pragma unroll
            for (int dr = 0; dr < 5; ++dr)
#pragma unroll
                for (int dc = 0; dc < 3; ++dc) d |= s_occ[lr + dr][lc + dc + 0] & 1;
        }
        s_dil[lr][lc] = d;
    }
    __syncthreads();
    bool far = false;
    for (int i = threadIdx.x; i < CF_TH * CF_TW; i += CF_THREADS) {
        const int lr = i / CF_TW, lc = i % CF_TW;
        const int r = r0 + lr, c = c0 + lc;
        if (r >= H || c >= W) continue;
        uint8_t e = 1;
#pragma unroll
        for (int dr = 0; dr < 5; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) e &= s_dil[lr + dr][lc + dc];
        const int sr = lr + CF_HR, sc = lc + CF_HC;
        double t = s_val[sr][sc];
        const bool one = (s_occ[sr][sc] & 2) != 0;
        bool filled = false;
        if (e && !one) {                                  // cl:41-43: closed == 255 and label != 1
            int neighbors = 0;
            double sum = 0.0;
#pragma unroll
            for (int dr = -2; dr <= 2; ++dr)
#pragma unroll
                for (int dc = -1; dc <= 1; ++dc)           // out-of-image neighbours were staged as "not 1"
                    if (s_occ[sr + dr][sc + dc] & 2) { neighbors += 1; sum = r3d::add(sum, s_val[sr + dr][sc + dc]); }
            if (neighbors > 0) t = __ddiv_rn(sum, (double)neighbors);        // cl:57
            filled = true;                                                   // cl:55,58: label = 1
        }
        const int64_t idx = base + (int64_t)r * W + c;
        out_train[idx] = t;
        if (out_label) out_label[idx] = filled ? 1.0 : in.lab(idx);
        if (closed_out) closed_out[idx] = e ? 255 : 0;
        far |= t > r3d::kEmptyRange;
    }
    if (far_flag && far) atomicOr(&far_flag[z], 1);
}

// one image per blockIdx.z, tiles on the (x, y) grid
template <class In>
__global__ void __launch_bounds__(CF_THREADS) k_close_fill(In in, int H, int W, int64_t img_stride,
                                                            double* __restrict__ out_train, double* __restrict__ out_label,
                                                            uint8_t* __restrict__ closed_out, int* __restrict__ far_flag) {
    const int z = blockIdx.z;
    close_fill_tile(in, H, W, (int64_t)z * img_stride, (int)blockIdx.y * CF_TH, (int)blockIdx.x * CF_TW, z, out_train, out_label,
                    closed_out, far_flag);
}

// the engine's variant: a persistent grid walks the (image, tile) tasks k_update listed for this round — after the
// first round only the few tiles around an inserted object change, so a grid over every tile of every image would be
// tens of thousands of CTAs that exit at once.  task = image * tiles_per_image + tile.
template <class In>
__global__ void __launch_bounds__(CF_THREADS) k_close_fill_tasks(In in, int H, int W, int64_t img_stride,
                                                                  double* __restrict__ out_train, int* __restrict__ far_flag,
                                                                  const int* __restrict__ tasks, const int* __restrict__ n_tasks) {
    const int tiles_x = (W + CF_TW - 1) / CF_TW, tiles = tiles_x * ((H + CF_TH - 1) / CF_TH);
    const int n = *n_tasks;
    for (int t = blockIdx.x; t < n; t += gridDim.x) {
        const int task = tasks[t], z = task / tiles, tile = task % tiles;
        close_fill_tile(in, H, W, (int64_t)z * img_stride, (tile / tiles_x) * CF_TH, (tile % tiles_x) * CF_TW, z, out_train,
                        (double*)nullptr, (uint8_t*)nullptr, far_flag);
        __syncthreads();                                 // the tile buffers are reused by the next task
    }
}
