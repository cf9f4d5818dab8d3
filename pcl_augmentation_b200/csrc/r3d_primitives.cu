// Primitive operators of the Real3D-Aug hot path on the reference's own float64 working-row layout (N x 9).
// These back the drop-in Python functions (fill_spherical, geometrical_front_view, smooth_out, cut_bounding_box).
#include "r3d_common.cuh"
#include "r3d_host.h"
#include "../../include/real3d_b200.h"
#include <cstring>

using namespace r3d;

// ------------------------------------------------------------------------------------ A1/A2 fill_spherical
__global__ void __launch_bounds__(256) k_fill_spherical_rows(double* __restrict__ rows, int64_t n,
                                                              unsigned long long* __restrict__ minmax_bits) {
    __shared__ unsigned long long s_min[8], s_max[8];
    unsigned long long lmin = R3D_EMPTY_U64, lmax = 0ull;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double* p = rows + i * 9;
        const double x = p[0], y = p[1], z = p[2];
        const double r = range3(x, y, z);
        const double az = azimuth(x, y);
        const double el = elevation(z, r);
        p[3] = r; p[4] = az; p[5] = el;
        const unsigned long long b = dbl_bits(el);
        lmin = min(lmin, b); lmax = max(lmax, b);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lmin; s_max[threadIdx.x >> 5] = lmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lmin = min(lmin, s_min[w]); lmax = max(lmax, s_max[w]); }
        atomicMax(&minmax_bits[0], lmax);
        atomicMin(&minmax_bits[1], lmin);
    }
}

__global__ void k_init_minmax(unsigned long long* mm) { mm[0] = 0ull; mm[1] = R3D_EMPTY_U64; }

extern "C" int r3d_fill_spherical(double* rows9, int64_t n, double* minmax_out, r3d_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || !minmax_out || (n > 0 && !rows9)) return r3d_fail(R3D_ERR_ARG, "r3d_fill_spherical: bad argument");
    unsigned long long* mm = reinterpret_cast<unsigned long long*>(minmax_out);
    k_init_minmax<<<1, 1, 0, stream>>>(mm); r3d_count_launch();
    if (n > 0) {
        int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
        k_fill_spherical_rows<<<grid, 256, 0, stream>>>(rows9, n, mm); r3d_count_launch();
    }
    return r3d_check_launch("r3d_fill_spherical");
}

// ----------------------------------------------------------------------------- A3 geometrical_front_view
__global__ void __launch_bounds__(256) k_zbuf_fill(unsigned long long* __restrict__ zbuf, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        zbuf[i] = R3D_EMPTY_U64;
}

__global__ void __launch_bounds__(256) k_project_rows(double* __restrict__ rows, int64_t n, ImageGeom g, int sample,
                                                       unsigned long long* __restrict__ zbuf, int* __restrict__ status) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double* p = rows + i * 9;
        const int row = bin_row(g, p[5]);
        const int col = bin_col(g, p[4]);
        const bool row_ok = row >= 0 && row < g.rows;
        if (!row_ok) {
            if (!sample) *status = R3D_ERR_ASSERT;          // od/ins:111
            continue;                                       // od/ins:108-109
        }
        if (!(col >= 0 && col < g.cols)) { *status = R3D_ERR_ASSERT; continue; }     // od/ins:113
        p[8] = (double)(row * g.pix_stride + col);                                  // od/ins:117,128
        atomicMin(&zbuf[row * g.cols + col], dbl_bits(p[3]));
    }
}

__global__ void __launch_bounds__(256) k_zbuf_to_images(const unsigned long long* __restrict__ zbuf, int64_t n,
                                                         double* __restrict__ train, double* __restrict__ label) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long b = zbuf[i];
        const bool hit = b != R3D_EMPTY_U64;
        train[i] = hit ? bits_dbl(b) : kEmptyRange;
        label[i] = hit ? 1.0 : -1.0;
    }
}

extern "C" int r3d_project_zbuffer(double* rows9, int64_t n, int num_row, int num_col, int pix_stride, double max_el,
                                   double min_el, int sample, double* train_out, double* label_out,
                                   uint64_t* zbuf_scratch, int* status_out, r3d_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_row <= 0 || num_col <= 0 || !train_out || !label_out || !zbuf_scratch || !status_out)
        return r3d_fail(R3D_ERR_ARG, "r3d_project_zbuffer: bad argument");
    const int64_t px = (int64_t)num_row * num_col;
    const int gpx = (int)std::min<int64_t>((px + 255) / 256, 148 * 8);
    cudaMemsetAsync(status_out, 0, sizeof(int), stream);
    k_zbuf_fill<<<gpx, 256, 0, stream>>>((unsigned long long*)zbuf_scratch, px); r3d_count_launch();
    if (n > 0) {
        ImageGeom g = make_geom(num_row, num_col, pix_stride, max_el, min_el);
        int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
        k_project_rows<<<grid, 256, 0, stream>>>(rows9, n, g, sample, (unsigned long long*)zbuf_scratch, status_out);
        r3d_count_launch();
    }
    k_zbuf_to_images<<<gpx, 256, 0, stream>>>((const unsigned long long*)zbuf_scratch, px, train_out, label_out);
    r3d_count_launch();
    return r3d_check_launch("r3d_project_zbuffer");
}

// ------------------------------------------------------------------------------------ A4 close + fill
// Input adaptor for the float64 train/label pair the reference passes to smooth_out.
struct F64Image {
    const double* train; const double* label;
    __device__ void load(int64_t i, double& v, uint8_t& o) const {
        const double l = label[i];
        v = train[i];
        o = (l > 0.0 ? 1 : 0) | (l == 1.0 ? 2 : 0);     // np.clip(label, 0, 1) -> 0/255 (cl:16-19); label == 1 (cl:41,48)
    }
    __device__ double lab(int64_t i) const { return label[i]; }
};

#include "r3d_closefill.cuh"

extern "C" int r3d_close_fill(const double* train_in, const double* label_in, int num_row, int num_col,
                              double* train_out, double* label_out, uint8_t* closed_out, r3d_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!train_in || !label_in || !train_out || !label_out || num_row <= 0 || num_col <= 0)
        return r3d_fail(R3D_ERR_ARG, "r3d_close_fill: bad argument");
    F64Image in{train_in, label_in};
    dim3 grid((num_col + CF_TW - 1) / CF_TW, (num_row + CF_TH - 1) / CF_TH, 1);
    k_close_fill<F64Image><<<grid, CF_THREADS, 0, stream>>>(in, num_row, num_col, 0, train_out, label_out, closed_out, nullptr);
    r3d_count_launch();
    return r3d_check_launch("r3d_close_fill");
}

// ------------------------------------------------------------------------------------ A8 cut_bounding_box
__global__ void __launch_bounds__(256) k_cut_box(const double* __restrict__ rows, int64_t n, int stride, Box box,
                                                  uint8_t* __restrict__ mask) {
    const BoxTest t = make_box_test(box);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* p = rows + i * stride;
        mask[i] = inside_box(t, p[0], p[1], p[2]) ? 1 : 0;
    }
}

extern "C" int r3d_cut_bounding_box(const double* rows, int64_t n, int row_stride, const double* box_host,
                                    uint8_t* mask_out, r3d_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || row_stride < 3 || !box_host || (n > 0 && (!rows || !mask_out)))
        return r3d_fail(R3D_ERR_ARG, "r3d_cut_bounding_box: bad argument");
    if (n == 0) return R3D_OK;
    Box b;
    b.cx = box_host[0]; b.cy = box_host[1]; b.cz = box_host[2];
    for (int i = 0; i < 9; ++i) b.m[i] = box_host[3 + i];
    b.length = box_host[12]; b.width = box_host[13]; b.height = box_host[14]; b.reach = box_host[15];
    int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    k_cut_box<<<grid, 256, 0, stream>>>(rows, n, row_stride, b, mask_out); r3d_count_launch();
    return r3d_check_launch("r3d_cut_bounding_box");
}

// --------------------------------------------------------- A5 / A7 helper: rotate about the sensor z-axis, shift z
// rotate_bounding_box (od/fs:97-102, ss/fs:72) and the z move of correct_height (od/fs:167) on rows of `stride` doubles
__global__ void __launch_bounds__(256) k_transform_rows(double* __restrict__ rows, int64_t n, int stride, double c, double s,
                                                         double dz) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double* p = rows + i * stride;
        const double x = p[0], y = p[1];
        p[0] = sub(mul(c, x), mul(s, y));
        p[1] = add(mul(s, x), mul(c, y));
        p[2] = add(p[2], dz);
    }
}

extern "C" int r3d_transform_points(double* rows, int64_t n, int row_stride, double cos_t, double sin_t, double dz,
                                    r3d_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || row_stride < 3 || (n > 0 && !rows)) return r3d_fail(R3D_ERR_ARG, "r3d_transform_points: bad argument");
    if (n == 0) return R3D_OK;
    int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    k_transform_rows<<<grid, 256, 0, stream>>>(rows, n, row_stride, cos_t, sin_t, dz); r3d_count_launch();
    return r3d_check_launch("r3d_transform_points");
}

// ------------------------------------------------------------------------------ A7 correct_height (single query)
struct LabelSet { int n; int v[R3D_MAX_SURFACE]; };
struct RadiiTable { double r2[R3D_NUM_RADII]; int ok[R3D_NUM_RADII]; };

template <int PASS>
__global__ void __launch_bounds__(256) k_road_level(const double* __restrict__ rows, int64_t n, int stride, int label_col,
                                                     LabelSet labels, double cx, double cy, double r2_limit,
                                                     unsigned long long* __restrict__ acc) {
    // acc[0] = min d2 bits, acc[1] = fixed-point z sum, acc[2] = count
    unsigned long long lmin = R3D_EMPTY_U64;
    long long zsum = 0; unsigned long long cnt = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* p = rows + i * stride;
        bool ok = false;
        for (int j = 0; j < labels.n; ++j) ok |= p[label_col] == (double)labels.v[j];
        if (!ok || !(p[2] > -3.0)) continue;
        const double dx = sub(p[0], cx), dy = sub(p[1], cy);
        const double d2 = add(mul(dx, dx), mul(dy, dy));
        if (PASS == 1) lmin = min(lmin, dbl_bits(d2));
        else if (d2 <= r2_limit) { zsum += __double2ll_rn(mul(p[2], 1099511627776.0)); ++cnt; }
    }
    if (PASS == 1) { if (lmin != R3D_EMPTY_U64) atomicMin(&acc[0], lmin); }
    else if (cnt) { atomicAdd(&acc[1], (unsigned long long)zsum); atomicAdd(&acc[2], cnt); }
}

extern "C" int r3d_road_level(const double* rows, int64_t n, int row_stride, int label_col, const int32_t* labels,
                              int n_labels, double cx, double cy, const double* radii_sq, const int32_t* radii_ok,
                              uint64_t* scratch3, double* level_out, int32_t* ok_out, r3d_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!level_out || !ok_out || !scratch3 || !radii_sq || !radii_ok || n_labels < 0 || n_labels > R3D_MAX_SURFACE ||
        row_stride <= label_col || label_col < 3)
        return r3d_fail(R3D_ERR_ARG, "r3d_road_level: bad argument");
    *ok_out = 0; *level_out = 0.0;
    if (n <= 0) return R3D_OK;
    LabelSet ls; ls.n = n_labels;
    for (int i = 0; i < n_labels; ++i) ls.v[i] = labels[i];
    unsigned long long h[3] = {R3D_EMPTY_U64, 0ull, 0ull};
    R3D_CUDA(cudaMemcpyAsync(scratch3, h, sizeof(h), cudaMemcpyHostToDevice, stream));
    int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    k_road_level<1><<<grid, 256, 0, stream>>>(rows, n, row_stride, label_col, ls, cx, cy, 0.0, (unsigned long long*)scratch3);
    r3d_count_launch();
    R3D_CUDA(cudaMemcpyAsync(h, scratch3, sizeof(h), cudaMemcpyDeviceToHost, stream));
    R3D_CUDA(cudaStreamSynchronize(stream));
    if (h[0] == R3D_EMPTY_U64) return R3D_OK;
    double d2; memcpy(&d2, &h[0], sizeof(double));
    int j = 0;
    while (j < R3D_NUM_RADII && !(d2 <= radii_sq[j])) ++j;                 // od/fs:152-160
    if (j >= R3D_NUM_RADII || !radii_ok[j]) return R3D_OK;
    k_road_level<2><<<grid, 256, 0, stream>>>(rows, n, row_stride, label_col, ls, cx, cy, radii_sq[j], (unsigned long long*)scratch3);
    r3d_count_launch();
    R3D_CUDA(cudaMemcpyAsync(h, scratch3, sizeof(h), cudaMemcpyDeviceToHost, stream));
    R3D_CUDA(cudaStreamSynchronize(stream));
    if (h[2] == 0) return R3D_OK;
    *level_out = ((double)(long long)h[1] / 1099511627776.0) / (double)h[2];    // od/fs:164 np.mean
    *ok_out = 1;
    return r3d_check_launch("r3d_road_level");
}

// --------------------------------------------------------------------------------- semseg addjust_map_2
struct Pose34 { double t[12]; };
__global__ void __launch_bounds__(256) k_adjust_map_rows(const double* __restrict__ rows, int64_t n, Pose34 T, double move_x,
                                                          double move_y, LabelSet ground, double* __restrict__ map, int sx,
                                                          int sy) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* p = rows + i * 9;
        if (!(p[2] < 1.5)) continue;                                   // ss/ins:207
        bool g = false;
        for (int j = 0; j < ground.n; ++j) g |= p[7] == (double)ground.v[j];
        if (g) continue;                                               // ss/ins:209-210
        const double wx = add(add(add(mul(T.t[0], p[0]), mul(T.t[1], p[1])), mul(T.t[2], p[2])), T.t[3]);
        const double wy = add(add(add(mul(T.t[4], p[0]), mul(T.t[5], p[1])), mul(T.t[6], p[2])), T.t[7]);
        const int ix = trunc_to_int(sub(wx, move_x)), iy = trunc_to_int(sub(wy, move_y));
        if (ix < 0 || iy < 0 || ix >= sx || iy >= sy) continue;        // reference: IndexError / negative wrap
        double* cell = map + (size_t)ix * sy + iy;
        if (*cell != 0.0) *cell = 4.0;                                 // ss/ins:221-222
    }
}

extern "C" int r3d_adjust_map(const double* rows9, int64_t n, const double* pose16_host, int64_t move_x, int64_t move_y,
                              const int32_t* ground_labels, int n_ground, double* map_dev, int size_x, int size_y,
                              r3d_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!pose16_host || !map_dev || size_x <= 0 || size_y <= 0 || n_ground < 0 || n_ground > R3D_MAX_SURFACE)
        return r3d_fail(R3D_ERR_ARG, "r3d_adjust_map: bad argument");
    if (n <= 0) return R3D_OK;
    Pose34 T; for (int i = 0; i < 12; ++i) T.t[i] = pose16_host[i];
    LabelSet g; g.n = n_ground; for (int i = 0; i < n_ground; ++i) g.v[i] = ground_labels[i];
    int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    k_adjust_map_rows<<<grid, 256, 0, stream>>>(rows9, n, T, (double)move_x, (double)move_y, g, map_dev, size_x, size_y);
    r3d_count_launch();
    return r3d_check_launch("r3d_adjust_map");
}
