// Cut-object database building: the point-in-box passes of
//   object_detection/cut_object/object_cut_out.py:90-168   (od/co)  expanded box + camera field of view + base box
//   semantic_segmentation/cut_object/cut_out.py:103-157    (ss/co)  box + class label
// batched over frames and boxes.  Both scripts call cut_bounding_box (cb:7-68) once or twice per ground-truth box, i.e.
// one N-length pass per box; here ONE pass over the points of a frame tests every box of the frame (boxes staged in
// shared memory, a float bounding-sphere test prunes, the exact fp64 cut_bounding_box test decides), and a stable
// chunked compaction writes the points of every box in their original order.
#include <algorithm>
#include "r3d_common.cuh"
#include "r3d_host.h"
#include "../../include/real3d_b200.h"

using namespace r3d;

namespace {

constexpr int CD_THREADS = 256;
constexpr int CD_CHUNK = 4096;            // points per CTA
constexpr int CD_MAX_BOXES = 64;          // boxes per frame (one bit each)
constexpr int CD_MAX_DROP = 8;

struct CutBox {
    BoxTest t;
    float cx, cy, cz, r2;                 // conservative bounding sphere (float, padded)
    int keep_label;                       // >= 0: emit only points with this label (ss/co:143); -1: any; -2: none
    int use_drop;                         // emit only points whose label is not in the drop list (od/co:150-152)
};

__device__ __forceinline__ CutBox make_cut_box(const Box& b, int keep_label, int use_drop) {
    CutBox c;
    c.t = make_box_test(b);
    // every point strictly inside lies within half the space diagonal of the box centre (z from the bottom)
    const double hz = 0.5 * b.height;
    const double rad = sqrt(0.25 * (b.length * b.length + b.width * b.width) + hz * hz);
    // the centre of a tilted box: bottom centre + (height / 2) along the box z-axis (column 2 of m)
    c.cx = (float)(b.cx + b.m[2] * hz); c.cy = (float)(b.cy + b.m[5] * hz); c.cz = (float)(b.cz + b.m[8] * hz);
    const float r = (float)rad * 1.0001f + 2e-3f;
    c.r2 = r * r;
    c.keep_label = keep_label; c.use_drop = use_drop;
    return c;
}

// camera field-of-view test of od/co:144 (cutout.py:73-122).  M1 = V2C.T @ R0.T (4 x 3) and P2 (3 x 4) arrive as the
// float32 values the reference holds, widened; numpy evaluates both products with a BLAS dgemm, i.e. per output element
// one accumulator walking k = 0..3 with fused multiply-adds (verified against numpy 2.3 / OpenBLAS 0.3.30).
struct Camera { double m1[12]; double p2[12]; double img_h, img_w; int enabled; };

__device__ __forceinline__ double dot4(double a0, double a1, double a2, double a3, double b0, double b1, double b2, double b3) {
    return __fma_rn(a3, b3, __fma_rn(a2, b2, __fma_rn(a1, b1, __dmul_rn(a0, b0))));
}
__device__ __forceinline__ bool in_fov(const Camera& c, double x, double y, double z) {
    const double r0 = dot4(x, y, z, 1.0, c.m1[0], c.m1[3], c.m1[6], c.m1[9]);        // lidar_to_rect (cutout.py:78-81)
    const double r1 = dot4(x, y, z, 1.0, c.m1[1], c.m1[4], c.m1[7], c.m1[10]);
    const double r2 = dot4(x, y, z, 1.0, c.m1[2], c.m1[5], c.m1[8], c.m1[11]);
    const double h0 = dot4(r0, r1, r2, 1.0, c.p2[0], c.p2[1], c.p2[2], c.p2[3]);      // rect_to_img (cutout.py:94-99)
    const double h1 = dot4(r0, r1, r2, 1.0, c.p2[4], c.p2[5], c.p2[6], c.p2[7]);
    const double h2 = dot4(r0, r1, r2, 1.0, c.p2[8], c.p2[9], c.p2[10], c.p2[11]);
    const double u = __ddiv_rn(h0, r2), v = __ddiv_rn(h1, r2);
    const double depth = __dsub_rn(h2, c.p2[11]);
    return u >= 0.0 && u < c.img_w && v >= 0.0 && v < c.img_h && depth >= 0.0;         // cutout.py:106-112
}

// coarse xy grid over +-81.92 m (5.12 m cells, the border cells also stand for everything beyond): cell -> bits of the
// boxes whose bounding sphere touches the cell, so a point looks at the zero to two boxes near it instead of all of them
constexpr int CD_GRID = 32;
constexpr float CD_GRID_HALF = 81.92f, CD_GRID_INV = 1.0f / 5.12f;
__device__ __forceinline__ int grid_cell(float v) {
    return (int)fminf(fmaxf((v + CD_GRID_HALF) * CD_GRID_INV, 0.0f), (float)(CD_GRID - 1));     // monotone; NaN -> 0
}

struct FrameCtx {
    CutBox box[CD_MAX_BOXES];
    unsigned long long grid[CD_GRID * CD_GRID];
    Camera cam;
    unsigned drop[CD_MAX_DROP];
    int n_boxes, n_drop;
};

__device__ void load_frame(FrameCtx& s, const double* __restrict__ boxes, const int* __restrict__ box_off,
                           const int* __restrict__ keep_label, const int* __restrict__ use_drop,
                           const double* __restrict__ cameras, const int* __restrict__ drop_labels, int n_drop, int f) {
    const int b0 = box_off[f], nb = min(box_off[f + 1] - b0, CD_MAX_BOXES);
    for (int i = threadIdx.x; i < CD_GRID * CD_GRID; i += blockDim.x) s.grid[i] = 0ull;
    __syncthreads();
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const double* r = boxes + (size_t)(b0 + j) * R3D_BOX_DOUBLES;
        Box b;
        b.cx = r[0]; b.cy = r[1]; b.cz = r[2];
        for (int i = 0; i < 9; ++i) b.m[i] = r[3 + i];
        b.length = r[12]; b.width = r[13]; b.height = r[14]; b.reach = r[15];
        const CutBox cb = make_cut_box(b, keep_label[b0 + j], use_drop[b0 + j]);
        s.box[j] = cb;
        const float rr = sqrtf(cb.r2) * 1.0001f + 1e-3f;
        const bool finite = rr < 1e30f && fabsf(cb.cx) < 1e30f && fabsf(cb.cy) < 1e30f;
        const int x0 = finite ? grid_cell(cb.cx - rr) : 0, x1 = finite ? grid_cell(cb.cx + rr) : CD_GRID - 1;
        const int y0 = finite ? grid_cell(cb.cy - rr) : 0, y1 = finite ? grid_cell(cb.cy + rr) : CD_GRID - 1;
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) atomicOr(&s.grid[y * CD_GRID + x], 1ull << j);
    }
    if (threadIdx.x < CD_MAX_DROP) s.drop[threadIdx.x] = threadIdx.x < n_drop ? (unsigned)drop_labels[threadIdx.x] : 0xFFFFFFFFu;
    if (threadIdx.x == 0) {
        s.n_boxes = nb; s.n_drop = n_drop;
        s.cam.enabled = 0;
        if (cameras) {
            const double* c = cameras + (size_t)f * 27;
            for (int i = 0; i < 12; ++i) { s.cam.m1[i] = c[i]; s.cam.p2[i] = c[12 + i]; }
            s.cam.img_h = c[24]; s.cam.img_w = c[25]; s.cam.enabled = c[26] != 0.0;
        }
    }
    __syncthreads();
}

// bits of the boxes of the frame that strictly contain the point (cut_bounding_box, cb:30-66)
__device__ __forceinline__ unsigned long long inside_bits(const FrameCtx& s, const float4& v) {
    unsigned long long m = 0ull;
    unsigned long long near = s.grid[grid_cell(v.y) * CD_GRID + grid_cell(v.x)];
    while (near) {
        const int j = __ffsll((long long)near) - 1; near &= near - 1;
        const CutBox& b = s.box[j];
        const float dx = v.x - b.cx, dy = v.y - b.cy, dz = v.z - b.cz;
        if (dx * dx + dy * dy + dz * dz > b.r2) continue;
        if (inside_box(b.t, (double)v.x, (double)v.y, (double)v.z)) m |= 1ull << j;
    }
    return m;
}
__device__ __forceinline__ bool emits(const FrameCtx& s, const CutBox& b, unsigned lab) {
    if (b.keep_label == -2) return false;                       // counted only (the expanded box of od/co:137-145)
    if (b.keep_label >= 0 && lab != (unsigned)b.keep_label) return false;
    if (b.use_drop)
        for (int i = 0; i < s.n_drop; ++i) if (lab == s.drop[i]) return false;
    return true;
}

// pass 1: per box the points inside, the points inside that the camera sees, and per (box, chunk) the points to emit
__global__ void __launch_bounds__(CD_THREADS) k_cut_count(const float4* __restrict__ xyzi, const unsigned* __restrict__ labels,
                                                          const long long* __restrict__ pt_off, int n_frames,
                                                          const double* __restrict__ boxes, const int* __restrict__ box_off,
                                                          const int* __restrict__ keep_label, const int* __restrict__ use_drop,
                                                          const double* __restrict__ cameras, const int* __restrict__ drop_labels,
                                                          int n_drop, int chunks, int* __restrict__ cnt_in,
                                                          int* __restrict__ cnt_fov, int* __restrict__ chunk_cnt) {
    const int f = blockIdx.y;
    if (f >= n_frames) return;
    __shared__ FrameCtx s;
    __shared__ int s_in[CD_MAX_BOXES], s_fov[CD_MAX_BOXES], s_emit[CD_MAX_BOXES];
    load_frame(s, boxes, box_off, keep_label, use_drop, cameras, drop_labels, n_drop, f);
    if (threadIdx.x < CD_MAX_BOXES) { s_in[threadIdx.x] = 0; s_fov[threadIdx.x] = 0; s_emit[threadIdx.x] = 0; }
    __syncthreads();
    const long long o = pt_off[f];
    const int n = (int)(pt_off[f + 1] - o);
    const int p0 = blockIdx.x * CD_CHUNK;
    for (int i = p0 + threadIdx.x; i < min(p0 + CD_CHUNK, n); i += CD_THREADS) {
        const float4 v = __ldg(&xyzi[o + i]);
        unsigned long long m = inside_bits(s, v);
        if (!m) continue;
        const unsigned lab = __ldg(&labels[o + i]);
        const bool fov = s.cam.enabled && in_fov(s.cam, (double)v.x, (double)v.y, (double)v.z);
        while (m) {
            const int j = __ffsll((long long)m) - 1; m &= m - 1;
            atomicAdd(&s_in[j], 1);
            if (fov) atomicAdd(&s_fov[j], 1);
            if (emits(s, s.box[j], lab)) atomicAdd(&s_emit[j], 1);
        }
    }
    __syncthreads();
    const int b0 = box_off[f];
    for (int j = threadIdx.x; j < s.n_boxes; j += CD_THREADS) {
        if (s_in[j]) atomicAdd(&cnt_in[b0 + j], s_in[j]);
        if (s_fov[j]) atomicAdd(&cnt_fov[b0 + j], s_fov[j]);
        chunk_cnt[(size_t)(b0 + j) * chunks + blockIdx.x] = s_emit[j];
    }
}

// per box: exclusive scan of its chunk counts (in place) and its total; then the box offsets into the packed output
__global__ void __launch_bounds__(1024) k_cut_offsets(int* __restrict__ chunk_cnt, int n_boxes, int chunks,
                                                      long long* __restrict__ out_off) {
    __shared__ long long s_tot[1024];
    __shared__ long long s_base;
    if (threadIdx.x == 0) { s_base = 0; out_off[0] = 0; }
    __syncthreads();
    for (int j0 = 0; j0 < n_boxes; j0 += 1024) {
        const int j = j0 + threadIdx.x;
        long long tot = 0;
        if (j < n_boxes) {
            int* c = chunk_cnt + (size_t)j * chunks;
            int run = 0;
            for (int k = 0; k < chunks; ++k) { const int v = c[k]; c[k] = run; run += v; }
            tot = run;
        }
        s_tot[threadIdx.x] = tot;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long run = s_base;
            for (int i = 0; i < min(1024, n_boxes - j0); ++i) { run += s_tot[i]; out_off[j0 + i + 1] = run; }
            s_base = run;
        }
        __syncthreads();
    }
}

// pass 2: the points of every box in their original order (x, y, z, intensity as read + label)
__global__ void __launch_bounds__(CD_THREADS) k_cut_write(const float4* __restrict__ xyzi, const unsigned* __restrict__ labels,
                                                          const long long* __restrict__ pt_off, int n_frames,
                                                          const double* __restrict__ boxes, const int* __restrict__ box_off,
                                                          const int* __restrict__ keep_label, const int* __restrict__ use_drop,
                                                          const int* __restrict__ drop_labels, int n_drop, int chunks,
                                                          const int* __restrict__ chunk_cnt, const long long* __restrict__ out_off,
                                                          float4* __restrict__ out_xyzi, unsigned* __restrict__ out_label,
                                                          int* __restrict__ out_index) {
    const int f = blockIdx.y;
    if (f >= n_frames) return;
    __shared__ FrameCtx s;
    __shared__ long long s_pos[CD_MAX_BOXES];
    __shared__ unsigned long long s_any;
    __shared__ int s_warp[CD_THREADS / 32];
    load_frame(s, boxes, box_off, keep_label, use_drop, nullptr, drop_labels, n_drop, f);
    const int b0 = box_off[f];
    for (int j = threadIdx.x; j < s.n_boxes; j += CD_THREADS)
        s_pos[j] = out_off[b0 + j] + chunk_cnt[(size_t)(b0 + j) * chunks + blockIdx.x];
    const long long o = pt_off[f];
    const int n = (int)(pt_off[f + 1] - o);
    const int p0 = blockIdx.x * CD_CHUNK, p1 = min(p0 + CD_CHUNK, n);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int t0 = p0; t0 < p1; t0 += CD_THREADS) {              // tiles in point order
        if (threadIdx.x == 0) s_any = 0ull;
        __syncthreads();
        const int i = t0 + threadIdx.x;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned lab = 0;
        unsigned long long m = 0ull;
        if (i < p1) {
            v = __ldg(&xyzi[o + i]);
            m = inside_bits(s, v);
            if (m) {
                lab = __ldg(&labels[o + i]);
                unsigned long long e = 0ull, q = m;
                while (q) { const int j = __ffsll((long long)q) - 1; q &= q - 1; if (emits(s, s.box[j], lab)) e |= 1ull << j; }
                m = e;
                if (m) atomicOr(&s_any, m);
            }
        }
        __syncthreads();
        unsigned long long any = s_any;
        __syncthreads();                                         // s_any is re-armed at the top of the next tile
        while (any) {                                            // usually zero or one box per tile
            const int j = __ffsll((long long)any) - 1; any &= any - 1;
            const bool hit = (m >> j) & 1ull;
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_warp[w] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int k = 0; k < CD_THREADS / 32; ++k) { if (k < w) before += s_warp[k]; total += s_warp[k]; }
            if (hit) {
                const long long pos = s_pos[j] + before + __popc(bal & ((1u << lane) - 1u));
                out_xyzi[pos] = v; out_label[pos] = lab;
                if (out_index) out_index[pos] = i;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_pos[j] += total;
            __syncthreads();
        }
    }
}

}  // namespace

extern "C" int r3d_cut_objects_count(const float* xyzi, const uint32_t* labels, const int64_t* point_offsets, int32_t n_frames,
                                     int32_t max_points, const double* boxes, const int32_t* box_offsets,
                                     const int32_t* keep_label, const int32_t* use_drop, int32_t n_boxes,
                                     int32_t max_boxes_per_frame, const double* cameras,
                                     const int32_t* drop_labels, int32_t n_drop, int32_t* count_inside, int32_t* count_fov,
                                     int32_t* chunk_counts, int64_t* out_offsets, r3d_stream stream_) {
    if (!xyzi || !labels || !point_offsets || !boxes || !box_offsets || !keep_label || !use_drop || !count_inside || !count_fov ||
        !chunk_counts || !out_offsets || n_frames <= 0 || n_boxes <= 0 || max_points < 0 || n_drop < 0 || n_drop > CD_MAX_DROP ||
        (n_drop > 0 && !drop_labels))
        return r3d_fail(R3D_ERR_ARG, "r3d_cut_objects_count: bad argument");
    if (max_boxes_per_frame < 0 || max_boxes_per_frame > CD_MAX_BOXES)
        return r3d_fail(R3D_ERR_CAPACITY, "r3d_cut_objects_count: at most 64 boxes per frame (split the frame)");
    cudaStream_t st = (cudaStream_t)stream_;
    const int chunks = std::max(1, (max_points + CD_CHUNK - 1) / CD_CHUNK);
    R3D_CUDA(cudaMemsetAsync(count_inside, 0, (size_t)n_boxes * sizeof(int), st));
    R3D_CUDA(cudaMemsetAsync(count_fov, 0, (size_t)n_boxes * sizeof(int), st));
    R3D_CUDA(cudaMemsetAsync(chunk_counts, 0, (size_t)n_boxes * chunks * sizeof(int), st));
    k_cut_count<<<dim3(chunks, n_frames), CD_THREADS, 0, st>>>((const float4*)xyzi, labels, (const long long*)point_offsets, n_frames,
                                                                boxes, box_offsets, keep_label, use_drop, cameras, drop_labels, n_drop,
                                                                chunks, count_inside, count_fov, chunk_counts);
    k_cut_offsets<<<1, 1024, 0, st>>>(chunk_counts, n_boxes, chunks, (long long*)out_offsets);
    r3d_count_launch(2);
    return r3d_check_launch("r3d_cut_objects_count");
}

extern "C" int r3d_cut_objects_write(const float* xyzi, const uint32_t* labels, const int64_t* point_offsets, int32_t n_frames,
                                     int32_t max_points, const double* boxes, const int32_t* box_offsets,
                                     const int32_t* keep_label, const int32_t* use_drop, int32_t n_boxes,
                                     const int32_t* drop_labels, int32_t n_drop, const int32_t* chunk_counts,
                                     const int64_t* out_offsets, float* out_xyzi, uint32_t* out_labels, int32_t* out_index,
                                     r3d_stream stream_) {
    if (!xyzi || !labels || !point_offsets || !boxes || !box_offsets || !keep_label || !use_drop || !chunk_counts || !out_offsets ||
        !out_xyzi || !out_labels || n_frames <= 0 || n_boxes <= 0 || max_points < 0 || n_drop < 0 || n_drop > CD_MAX_DROP)
        return r3d_fail(R3D_ERR_ARG, "r3d_cut_objects_write: bad argument");
    cudaStream_t st = (cudaStream_t)stream_;
    const int chunks = std::max(1, (max_points + CD_CHUNK - 1) / CD_CHUNK);
    k_cut_write<<<dim3(chunks, n_frames), CD_THREADS, 0, st>>>((const float4*)xyzi, labels, (const long long*)point_offsets, n_frames,
                                                                boxes, box_offsets, keep_label, use_drop, drop_labels, n_drop, chunks,
                                                                chunk_counts, (const long long*)out_offsets, (float4*)out_xyzi,
                                                                out_labels, out_index);
    r3d_count_launch();
    return r3d_check_launch("r3d_cut_objects_write");
}
