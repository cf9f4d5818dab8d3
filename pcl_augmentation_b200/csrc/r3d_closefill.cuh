// A4: class_closing + smooth_out (cl:9-62) as one tiled stencil kernel.
//   closing(rectangle(5, 3)) = 5 rows x 3 cols dilate then erode on the binary occupancy, pixels outside the image
//   ignored, no azimuth wrap; pixels switched on by the closing that hold no point get the mean of their occupied
//   neighbours' ORIGINAL ranges, summed in the reference's (drow, dcol) order in fp64 (cl:46-51,57).
// Templated on the input adaptor so the same code serves the float64 train/label pair of the drop-in smooth_out and
// the engine's raw uint64 z-buffer.  HBM traffic: one read of the input tile (+halo from L2) and one write of the
// output -> 16 B/pixel for the engine variant.
#pragma once
#include "r3d_common.cuh"

constexpr int CF_TH = 16;       // output rows per CTA
constexpr int CF_TW = 64;       // output cols per CTA
constexpr int CF_THREADS = 256;

template <class In>
__global__ void __launch_bounds__(CF_THREADS) k_close_fill(In in, int H, int W, int64_t img_stride,
                                                            double* __restrict__ out_train, double* __restrict__ out_label,
                                                            uint8_t* __restrict__ closed_out, int* __restrict__ far_flag,
                                                            const int* __restrict__ gate) {
    const int z = blockIdx.z;
    if (gate && !gate[z]) return;
    __shared__ uint8_t s_occ[CF_TH + 8][CF_TW + 4];
    __shared__ uint8_t s_dil[CF_TH + 4][CF_TW + 2];
    const int r0 = blockIdx.y * CF_TH, c0 = blockIdx.x * CF_TW;
    const int64_t base = (int64_t)z * img_stride;
    for (int i = threadIdx.x; i < (CF_TH + 8) * (CF_TW + 4); i += CF_THREADS) {
        const int lr = i / (CF_TW + 4), lc = i % (CF_TW + 4);
        const int r = r0 - 4 + lr, c = c0 - 2 + lc;
        s_occ[lr][lc] = (r >= 0 && r < H && c >= 0 && c < W) ? (in.occ(base + (int64_t)r * W + c) ? 1 : 0) : 0;
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
                for (int dc = 0; dc < 3; ++dc) d |= s_occ[lr + dr][lc + dc];
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
        const int64_t idx = base + (int64_t)r * W + c;
        double t = in.val(idx);
        bool one = in.is_one(idx);
        double lab = in.lab(idx);
        if (e && !one) {                                  // cl:41-43: closed == 255 and label != 1
            int neighbors = 0;
            double sum = 0.0;
            for (int dr = -2; dr <= 2; ++dr)
                for (int dc = -1; dc <= 1; ++dc) {
                    const int rr = r + dr, cc = c + dc;
                    if (rr >= 0 && rr < H && cc >= 0 && cc < W) {
                        const int64_t j = base + (int64_t)rr * W + cc;
                        if (in.is_one(j)) { neighbors += 1; sum = r3d::add(sum, in.val(j)); }
                    }
                }
            if (neighbors > 0) t = __ddiv_rn(sum, (double)neighbors);        // cl:57
            lab = 1.0;                                                       // cl:55,58
        }
        out_train[idx] = t;
        if (out_label) out_label[idx] = lab;
        if (closed_out) closed_out[idx] = e ? 255 : 0;
        far |= t > r3d::kEmptyRange;
    }
    if (far_flag && far) atomicOr(&far_flag[z], 1);
}
