// Host side of the batched Real3D-Aug engine: device memory, H2D/D2H staging, the per-round launch sequence and
// the C ABI declared in include/real3d_b200.h.
#include "r3d_engine_kernels.cuh"
#include "r3d_host.h"

#include <cmath>
#include <chrono>
#include <cstdlib>
#include <thread>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace r3d;

#define R3D_MAX_SUB 16

namespace {

enum KernelId {
    KID_INGEST, KID_CTRL, KID_UPDATE, KID_CLEAR, KID_PROJECT, KID_CLOSEFILL, KID_ADJUST, KID_ONMAP, KID_HEIGHT, KID_COLLIDE,
    KID_GRID, KID_OCCL, KID_SELECT, KID_OUT, KID_MINMAX, KID_PROJECT0, KID_CLOSEFILL0, KID_CLEAR0, KID_MINMAX0, KID_WALK, KID_PREP, KID_SCATTER,
    KID_COUNT
};
const char* kKernelNames[KID_COUNT] = {
    "ingest_spherical", "ctrl", "update_mask_patch", "clear_images", "project_zbuffer", "close_fill", "adjust_map",
    "onmap", "road_level", "collide", "index_build", "occlusion_count", "select_emit", "compact_output", "minmax_elevation",
    // round 0 of a run re-projects every scan in full; later rounds only the scans whose elevation range moved
    "project_zbuffer_full", "close_fill_full", "clear_images_full", "minmax_elevation_full",
    // the per-scan persistent walker (one CTA per scan runs all the slots / tries of its scan) and its set-up
    "scan_walk", "walk_prepare",
    // fused streaming passes: ingest_spherical = A1 + A2 + index counts + z-buffer clear, index_build = scans + distance
    // transform, scatter_project = CSR scatter of the three indices + A3 (pix ids, z-buffer)
    "scatter_project"};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return R3D_OK;
        cudaError_t err = cudaMalloc((void**)&p, count * sizeof(T));
        if (err != cudaSuccess) { p = nullptr; return r3d_fail_cuda(err, "cudaMalloc"); }
        return R3D_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

}  // namespace

struct r3d_engine {
    r3d_engine_cfg cfg;
    EngineDev dev;
    cudaStream_t stream = nullptr;
    int n_scans = 0;
    int max_n0 = 0;
    int n_sms = 148;
    int device = 0;                              // every API call binds the calling thread to the engine's device first
    int n_sub = 4;                               // sub-batches advanced concurrently, each on its own stream
    cudaStream_t sub_stream[R3D_MAX_SUB] = {nullptr};
    cudaEvent_t sub_done[R3D_MAX_SUB] = {nullptr};
    cudaEvent_t ev_armed = nullptr;
    cudaEvent_t ev_wait = nullptr;               // blocking-sync event for host waits on the engine stream
    // state of a run in progress (r3d_engine_run_until may return before the batch is finished)
    struct Sub { int b0, n, round; bool done; long long left; unsigned seq_base; EngineDev d; cudaStream_t st; };
    Sub sub[R3D_MAX_SUB];
    // one round of a sub-batch (k_ctrl ... k_select_emit) captured as a CUDA graph: every argument is constant between
    // rounds (the round number lives on the device), so a round costs the host ONE launch instead of twelve
    struct RoundGraph { cudaGraphExec_t exec = nullptr; EngineDev d; int ns = 0, chunks_all = 0, task_ctas = 0, sel_pts = 0, kernels = 0; };
    RoundGraph round_graph[R3D_MAX_SUB];
    bool use_graphs = true, capturing = false;
    bool fresh = false;                          // ingest just ran: z-buffer / pix ids / elevation range of the originals are valid
    bool walked = false;                         // the last run used the walker (its step count is in h_offsets)
    bool staged = false;                         // true: the staged round kernels (debug / probe path); false: the per-scan walker
    bool run_active = false;
    int run_nsub = 0, run_done = 0, run_rounds = 0;
    bool objects_set = false, yaw_set = false, batch_loaded = false, ran = false;
    int last_rounds = 0;
    // device buffers
    DevBuf<float4> xyzi, out_xyzi, gpts, apts;
    DevBuf<unsigned long long> sel_keys;
    DevBuf<double> sel_r, tail_x, tail_y, tail_z, r, el, smooth, poses, obj_x, obj_y, obj_z, cos_k, sin_k, radii_sq, cand_level,
        inserted_box;
    DevBuf<float> tail_i, obj_i, check, out_check;
    DevBuf<unsigned> label, dmask, vmask, occ_win, unplaceable, obj_label, out_label, tickets;
    DevBuf<unsigned short> col, cand_list, label16, out_label16;
    DevBuf<unsigned char> label1;
    DevBuf<long long> pt_off;
    DevBuf<int> chunk_cnt, pix, gate_update, gate_try, gate_apply, gate_full, gate_patch, cf_rect, col_off, col_idx, acell, active_count, far_arr, od_map_dims, counts, perms, class_list_off,
        class_list, radii_ok, cand_v, gcell, n_list, feas, occ_pix, sel_pix, inserted, n0_arr, nbox0_arr, work_cnt, full_list, cf_tasks, need2, occ_far;
    DevBuf<unsigned> occ_cnt;
    DevBuf<unsigned> round_ctl;
    DevBuf<unsigned char> alive, od_maps, ss_map, cand_flags, gnear, gscratch;
    DevBuf<unsigned long long> zraw, obj_raw, stats;
    DevBuf<long long> od_map_off, out_count, out_off, check_off;
    DevBuf<ScanState> st;
    DevBuf<Box> boxes, boxes0;                   // boxes0: the scene boxes as loaded (re-arm drops the inserted ones, device to device)
    DevBuf<BoxTest> box_tests;
    DevBuf<ObjBox> obj, try_obj;
    DevBuf<ClassCfg> classes;
    CUtensorMap zraw_tmap;            // [B][H][W] u64 z-buffer as a 3-D tiled tensor, box = one staged close/fill tile
    bool zraw_tmap_ok = false;
    // host copies needed for re-arming
    std::vector<Box> h_boxes;
    std::vector<int> h_nbox0;
    unsigned long long* h_words = nullptr;   // mapped pinned: one word per round slot, written by k_ctrl
    unsigned long long* d_words = nullptr;   // device alias of h_words
    unsigned seq = 0;
    long long* h_offsets = nullptr;     // pinned, 2 * (max_scans + 1)
    // profiling
    bool profile = false;
    double prof_ms[KID_COUNT] = {0};
    int64_t prof_launches[KID_COUNT] = {0};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending_events;
    std::vector<cudaEvent_t> event_pool;
};

namespace {

__global__ void k_box_tests(const Box* boxes, BoxTest* tests, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tests[i] = make_box_test(boxes[i]);
}
// packed 16-bit labels (contiguous, per-scan offsets) -> the engine's u32 label rows
__global__ void k_widen_labels(const unsigned short* src, const long long* pt_off, unsigned* label, int P, int n_scans) {
    const int b = blockIdx.y;
    if (b >= n_scans) return;
    const long long o = pt_off[b] - pt_off[0];
    const int cnt = (int)(pt_off[b + 1] - pt_off[b]);
    const int p0 = blockIdx.x * CHUNK;
    for (int p = p0 + threadIdx.x; p < min(p0 + CHUNK, cnt); p += blockDim.x) label[(size_t)b * P + p] = src[o + p];
}
// one "is Road" bit per point (object detection) -> {road_label, some other label}
__global__ void k_expand_label_bits(const unsigned char* bits, const long long* pt_off, unsigned* label, int P, int n_scans,
                                    unsigned road_label) {
    const int b = blockIdx.y;
    if (b >= n_scans) return;
    const long long o = pt_off[b] - pt_off[0];
    const int cnt = (int)(pt_off[b + 1] - pt_off[b]);
    const int p0 = blockIdx.x * CHUNK;
    const unsigned other = road_label == 1u ? 2u : 1u;       // od/ins:353-355 relabels every non-Road point to 1
    for (int p = p0 + threadIdx.x; p < min(p0 + CHUNK, cnt); p += blockDim.x) {
        const long long i = o + p;
        label[(size_t)b * P + p] = ((bits[i >> 3] >> (i & 7)) & 1u) ? road_label : other;
    }
}
__global__ void k_narrow_labels(const unsigned* src, unsigned short* dst, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = (unsigned short)src[i];
}
__global__ void k_set_round(unsigned* round_ctl, unsigned seq_base) { round_ctl[0] = 0u; round_ctl[1] = seq_base; }
__global__ void k_fill_u64(unsigned long long* p, size_t n, unsigned long long v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

struct Launcher {          // wraps every launch: launch counter + optional CUDA-event timing on the launching stream
    r3d_engine* eng;
    int kid;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    Launcher(r3d_engine* e, int k, cudaStream_t stream = nullptr) : eng(e), kid(k), st(stream ? stream : e->stream) {
        if (eng->profile) {
            a = get_event(); b = get_event();
            cudaEventRecord(a, st);
        }
    }
    cudaEvent_t get_event() {
        if (!eng->event_pool.empty()) { cudaEvent_t ev = eng->event_pool.back(); eng->event_pool.pop_back(); return ev; }
        cudaEvent_t ev; cudaEventCreate(&ev); return ev;
    }
    ~Launcher() {
        if (eng->capturing) return;          // counted per graph launch
        r3d_count_launch();
        eng->prof_launches[kid] += 1;
        if (eng->profile) {
            cudaEventRecord(b, st);
            eng->pending_events.push_back({kid, {a, b}});
        }
    }
};

void drain_events(r3d_engine* eng) {
    for (auto& pe : eng->pending_events) {
        float ms = 0.f;
        cudaEventSynchronize(pe.second.second);
        cudaEventElapsedTime(&ms, pe.second.first, pe.second.second);
        eng->prof_ms[pe.first] += ms;
        eng->event_pool.push_back(pe.second.first);
        eng->event_pool.push_back(pe.second.second);
    }
    eng->pending_events.clear();
}

size_t occl_smem_bytes(const EngineDev& d) { return (size_t)d.dwords * sizeof(unsigned) + (size_t)(d.K + 2) * sizeof(unsigned short); }
constexpr int SEL_SMEM_PTS = 4096;       // object points k_select_emit keeps in shared memory; larger objects use global scratch
int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
// dynamic shared memory of k_select_emit: sort keys, ranges, object tile, pixel ids, dilation + visibility bits
size_t select_smem_bytes(int max_pts) {
    return (size_t)next_pow2(max_pts) * 8 + (size_t)max_pts * 8 + (size_t)SEL_TILE_PX * 8 + (size_t)max_pts * 4 +
           2 * (size_t)(SEL_TILE_PX / 32) * 4 + 16;
}

}  // namespace

#define TRY(x) do { int _rc = (x); if (_rc != R3D_OK) return _rc; } while (0)

// Wait for the engine stream WITHOUT spinning: the waits below cover millisecond-long transfers and whole runs, and a
// rank keeps several engine threads; cudaStreamSynchronize would burn a host core per waiting thread (8 ranks x 4
// pipelined engines on a 32-vCPU box starve the threads that launch the rounds).  A blocking-sync event sleeps instead.
static cudaError_t engine_wait(r3d_engine* eng) {
    if (!eng->ev_wait) {
        cudaError_t e = cudaEventCreateWithFlags(&eng->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(eng->ev_wait, eng->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(eng->ev_wait);
}

// TMA descriptor of the z-buffer for k_close_fill_tma (r3d_closefill.cuh).  cuTensorMapEncodeTiled is a driver API entry:
// fetched through the runtime so that the library does not link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_zraw_tensor_map(CUtensorMap* tm, void* base, int W, int H, int B) {
    if (getenv("R3D_NO_TMA") || (W & 1) || ((uintptr_t)base & 15)) return false;       // row pitch must be a multiple of 16 bytes
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 8, (cuuint64_t)W * H * 8};
    const cuuint32_t box[3] = {(cuuint32_t)CF_SW, (cuuint32_t)CF_SH, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return ((EncodeTiledFn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int r3d_engine_create(const r3d_engine_cfg* cfg, r3d_engine** out) {
    if (!cfg || !out) return r3d_fail(R3D_ERR_ARG, "r3d_engine_create: null argument");
    if (cfg->rows <= 0 || cfg->cols <= 0 || cfg->cols > 65535 || cfg->yaw_steps <= 0 || cfg->n_classes <= 0 ||
        cfg->n_classes > R3D_MAX_CLASSES || cfg->max_scans <= 0 || cfg->max_points <= 0 || cfg->max_inserted <= 0 ||
        cfg->max_tries <= 0 || cfg->max_events <= 0 || cfg->max_boxes <= 0 || (cfg->task != 0 && cfg->task != 1))
        return r3d_fail(R3D_ERR_ARG, "r3d_engine_create: bad configuration");
    if (cfg->task == 1 && (cfg->map_window <= 0 || cfg->map_window % 32 != 0))
        return r3d_fail(R3D_ERR_ARG, "r3d_engine_create: map_window must be a positive multiple of 32");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return r3d_fail(R3D_ERR_CUDA, "r3d_engine_create: no CUDA device (this library has no CPU fallback)");
    r3d_engine* eng = new r3d_engine();
    eng->cfg = *cfg;
    R3D_CUDA(cudaGetDevice(&eng->device));
    // the round kernels (short, latency bound) run on high-priority streams; uploads, the one-off spherical ingest /
    // index builds and the output compaction of OTHER engines on the same GPU (ScanPipeline) must not sit in front of them
    int prio_least = 0, prio_greatest = 0;
    R3D_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    R3D_CUDA(cudaStreamCreateWithPriority(&eng->stream, cudaStreamNonBlocking, prio_least));
    eng->n_sub = (cfg->flags >> 8) & 31;
    eng->use_graphs = !(cfg->flags & 2);
    eng->staged = (cfg->flags & 8) != 0;
    if (const char* env = getenv("R3D_STAGED")) eng->staged = atoi(env) != 0;
    if (const char* env = getenv("R3D_GRAPHS")) eng->use_graphs = atoi(env) != 0;
    if (const char* env = getenv("R3D_SUBBATCHES")) eng->n_sub = atoi(env);
    if (eng->n_sub <= 0) eng->n_sub = 4;
    eng->n_sub = std::min(eng->n_sub, R3D_MAX_SUB);
    for (int i = 0; i < R3D_MAX_SUB; ++i) {
        R3D_CUDA(cudaStreamCreateWithPriority(&eng->sub_stream[i], cudaStreamNonBlocking, prio_greatest));
        R3D_CUDA(cudaEventCreateWithFlags(&eng->sub_done[i], cudaEventDisableTiming));
    }
    R3D_CUDA(cudaEventCreateWithFlags(&eng->ev_armed, cudaEventDisableTiming));
    EngineDev& d = eng->dev;
    memset(&d, 0, sizeof(d));
    d.task = cfg->task; d.rows = cfg->rows; d.cols = cfg->cols; d.hw = cfg->rows * cfg->cols; d.K = cfg->yaw_steps;
    d.max_tries = cfg->max_tries; d.n_classes = cfg->n_classes; d.B = cfg->max_scans; d.max_points = cfg->max_points;
    d.max_inserted = cfg->max_inserted + (16 - (cfg->max_points + cfg->max_inserted) % 16) % 16;   // rows of P: 16-aligned
    d.P = cfg->max_points + d.max_inserted; d.max_boxes = cfg->max_boxes;
    d.max_events = cfg->max_events; d.road_label = cfg->road_label; d.n_road_indexes = cfg->n_road_indexes;
    d.map_window = cfg->task == 1 ? cfg->map_window : 32; d.dwords = (d.hw + 31) / 32;
    for (int i = 0; i < R3D_MAX_SURFACE; ++i) d.road_indexes[i] = cfg->road_indexes[i];
    d.step_rad = 2.0 * 3.14159265358979323846 / (double)cfg->yaw_steps;
    d.grid_cell = cfg->grid_cell > 0 ? cfg->grid_cell : 0.5;
    d.grid_inv_cell = (float)(1.0 / d.grid_cell);
    d.G = 2 * (cfg->grid_half > 0 ? cfg->grid_half : 200);
    d.force_full = (cfg->flags & 1) ? 1 : 0;
    const size_t B = d.B, P = d.P, K1 = d.K + 1, HW = d.hw;
    TRY(eng->xyzi.alloc(B * d.max_points)); TRY(eng->tail_x.alloc(B * d.max_inserted)); TRY(eng->tail_y.alloc(B * d.max_inserted));
    TRY(eng->tail_z.alloc(B * d.max_inserted)); TRY(eng->tail_i.alloc(B * d.max_inserted)); TRY(eng->label.alloc(B * P));
    TRY(eng->r.alloc(B * P)); TRY(eng->el.alloc(B * P)); TRY(eng->col.alloc(B * P)); TRY(eng->pix.alloc(B * P));
    TRY(eng->alive.alloc(B * P)); TRY(eng->zraw.alloc(B * HW)); TRY(eng->obj_raw.alloc(B * HW)); TRY(eng->smooth.alloc(B * HW));
    eng->zraw_tmap_ok = make_zraw_tensor_map(&eng->zraw_tmap, eng->zraw.p, d.cols, d.rows, (int)B);
    TRY(eng->dmask.alloc(B * d.dwords)); TRY(eng->vmask.alloc(B * d.dwords)); TRY(eng->st.alloc(B));
    TRY(eng->gate_update.alloc(B)); TRY(eng->gate_try.alloc(B)); TRY(eng->gate_apply.alloc(B)); TRY(eng->gate_full.alloc(B));
    TRY(eng->gate_patch.alloc(B)); TRY(eng->cf_rect.alloc(B * 4)); TRY(eng->active_count.alloc(R3D_MAX_SUB * 64 * 2));
    TRY(eng->round_ctl.alloc(R3D_MAX_SUB * 32));
    d.cf_tiles_x = (d.cols + CF_TW - 1) / CF_TW;
    d.cf_tiles = d.cf_tiles_x * ((d.rows + CF_TH - 1) / CF_TH);
    TRY(eng->work_cnt.alloc(B * 4)); TRY(eng->full_list.alloc(B)); TRY(eng->cf_tasks.alloc(B * (size_t)d.cf_tiles));
    R3D_CUDA(cudaMemset(eng->work_cnt.p, 0, B * 4 * sizeof(int)));
    TRY(eng->need2.alloc(B));
    R3D_CUDA(cudaMemset(eng->need2.p, 0, B * sizeof(int)));
    d.cand_window = (cfg->flags & 4) ? 0 : 48;
    if (const char* env = getenv("R3D_WINDOW")) d.cand_window = std::max(0, atoi(env));
    if (d.task != 0) d.cand_window = 0;          // semseg walks the yaws with a carried z shift (ss/fs:146-147): no window
    TRY(eng->far_arr.alloc(B)); TRY(eng->boxes.alloc(B * d.max_boxes)); TRY(eng->box_tests.alloc(B * d.max_boxes));
    TRY(eng->boxes0.alloc(B * d.max_boxes));
    TRY(eng->poses.alloc(B * 16)); TRY(eng->occ_win.alloc(B * ((size_t)d.map_window * d.map_window / 32)));
    TRY(eng->counts.alloc(B * d.n_classes)); TRY(eng->cos_k.alloc(K1)); TRY(eng->sin_k.alloc(K1));
    TRY(eng->radii_sq.alloc(R3D_NUM_RADII)); TRY(eng->radii_ok.alloc(R3D_NUM_RADII)); TRY(eng->classes.alloc(R3D_MAX_CLASSES));
    TRY(eng->cand_flags.alloc(B * K1)); TRY(eng->cand_level.alloc(B * K1)); TRY(eng->cand_v.alloc(B * K1));
    TRY(eng->cand_list.alloc(B * K1)); TRY(eng->n_list.alloc(B)); TRY(eng->tickets.alloc(B * 4));
    R3D_CUDA(cudaMemset(eng->tickets.p, 0, B * 4 * sizeof(unsigned)));
    TRY(eng->feas.alloc(B * d.K)); TRY(eng->inserted.alloc(B * d.max_events * 4)); TRY(eng->inserted_box.alloc(B * d.max_events * 8));
    TRY(eng->check.alloc(B * d.max_inserted * 5)); TRY(eng->out_count.alloc(B)); TRY(eng->out_off.alloc(B + 1));
    TRY(eng->check_off.alloc(B + 1)); TRY(eng->out_xyzi.alloc(B * P)); TRY(eng->out_label.alloc(B * P));
    TRY(eng->out_check.alloc(B * d.max_inserted * 5)); TRY(eng->n0_arr.alloc(B)); TRY(eng->nbox0_arr.alloc(B));
    TRY(eng->gcell.alloc(B * (size_t)d.G * d.G)); TRY(eng->gpts.alloc(B * d.max_points));
    TRY(eng->gnear.alloc(B * (size_t)d.G * d.G)); TRY(eng->gscratch.alloc(B * (size_t)d.G * d.G));
    TRY(eng->col_off.alloc(B * (size_t)(d.cols + 1))); TRY(eng->col_idx.alloc(B * d.max_points));
    TRY(eng->acell.alloc(B * (size_t)d.G * d.G)); TRY(eng->apts.alloc(B * d.max_points));
    TRY(eng->try_obj.alloc(B));
    d.max_chunks = (d.P + CHUNK - 1) / CHUNK;
    TRY(eng->chunk_cnt.alloc(B * (size_t)d.max_chunks));
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&eng->n_sms, cudaDevAttrMultiProcessorCount, dev); }
    TRY(eng->od_map_off.alloc(B * 2 + 1)); TRY(eng->od_map_dims.alloc(B * 8)); TRY(eng->stats.alloc(R3D_N_STATS));
    R3D_CUDA(cudaMemset(eng->stats.p, 0, R3D_N_STATS * sizeof(unsigned long long)));
    TRY(eng->occ_far.alloc(B * (OCC_FAR_CAP + 1)));
    if (cfg->task == 1) TRY(eng->occ_cnt.alloc(B * (size_t)d.map_window * d.map_window));
    R3D_CUDA(cudaMemset(eng->occ_far.p, 0xFF, B * (OCC_FAR_CAP + 1) * sizeof(int)));
    R3D_CUDA(cudaHostAlloc((void**)&eng->h_words, R3D_MAX_SUB * 64 * sizeof(unsigned long long), cudaHostAllocMapped));
    memset(eng->h_words, 0, R3D_MAX_SUB * 64 * sizeof(unsigned long long));
    R3D_CUDA(cudaHostGetDevicePointer((void**)&eng->d_words, eng->h_words, 0));
    R3D_CUDA(cudaMallocHost((void**)&eng->h_offsets, (2 * (B + 1) + 1) * sizeof(long long)));
    std::vector<ClassCfg> cls(R3D_MAX_CLASSES);
    for (int c = 0; c < R3D_MAX_CLASSES; ++c) {
        const r3d_class_cfg& s = cfg->classes[c];
        cls[c].min_points = s.min_points; cls[c].map_sel = s.map_sel; cls[c].map_ok_mask = s.map_ok_mask;
        cls[c].pedestrian = s.pedestrian; cls[c].n_surface = std::min(std::max(s.n_surface, 0), R3D_MAX_SURFACE);
        for (int i = 0; i < R3D_MAX_SURFACE; ++i) cls[c].surface[i] = s.surface[i];
        cls[c].ok_slots = 0u;
    }
    d.n_surf_all = 0;
    for (int c = 0; c < cfg->n_classes; ++c)
        for (int i = 0; i < cls[c].n_surface; ++i) {
            int slot = -1;
            for (int j = 0; j < d.n_surf_all; ++j) if (d.surf_all[j] == cls[c].surface[i]) slot = j;
            if (slot < 0) {
                if (d.n_surf_all >= 15) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_create: more than 15 distinct surface labels");
                slot = d.n_surf_all; d.surf_all[d.n_surf_all++] = cls[c].surface[i];
            }
            cls[c].ok_slots |= 1u << slot;
        }
    if (d.max_points + d.max_inserted > (1 << APT_IDX_BITS)) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_create: more than 2^27 points per scan");
    R3D_CUDA(cudaMemcpy(eng->classes.p, cls.data(), cls.size() * sizeof(ClassCfg), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->radii_sq.p, cfg->radii_sq, sizeof(cfg->radii_sq), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->radii_ok.p, cfg->radii_ok, sizeof(cfg->radii_ok), cudaMemcpyHostToDevice));
    k_fill_u64<<<148 * 4, 256, 0, eng->stream>>>(eng->obj_raw.p, B * HW, R3D_EMPTY_U64); r3d_count_launch();
    R3D_CUDA(cudaMemsetAsync(eng->dmask.p, 0, B * d.dwords * sizeof(unsigned), eng->stream));
    R3D_CUDA(cudaStreamSynchronize(eng->stream));
    // wire the device view
    d.xyzi = eng->xyzi.p; d.tail_x = eng->tail_x.p; d.tail_y = eng->tail_y.p; d.tail_z = eng->tail_z.p; d.tail_i = eng->tail_i.p;
    d.label = eng->label.p; d.r = eng->r.p; d.el = eng->el.p; d.col = eng->col.p; d.pix = eng->pix.p; d.alive = eng->alive.p;
    d.zraw = eng->zraw.p; d.obj_raw = eng->obj_raw.p; d.smooth = eng->smooth.p; d.dmask = eng->dmask.p; d.vmask = eng->vmask.p;
    d.st = eng->st.p; d.gate_update = eng->gate_update.p; d.gate_try = eng->gate_try.p; d.gate_apply = eng->gate_apply.p;
    d.gate_full = eng->gate_full.p; d.gate_patch = eng->gate_patch.p; d.cf_rect = eng->cf_rect.p;
    d.work_cnt = eng->work_cnt.p; d.full_list = eng->full_list.p; d.cf_tasks = eng->cf_tasks.p; d.need2 = eng->need2.p;
    d.active_count = eng->active_count.p; d.far_arr = eng->far_arr.p; d.boxes = eng->boxes.p; d.box_tests = eng->box_tests.p;
    d.poses = eng->poses.p; d.occ_win = eng->occ_win.p; d.counts = eng->counts.p; d.cos_k = eng->cos_k.p; d.sin_k = eng->sin_k.p;
    d.radii_sq = eng->radii_sq.p; d.radii_ok = eng->radii_ok.p; d.classes = eng->classes.p; d.cand_flags = eng->cand_flags.p;
    d.cand_level = eng->cand_level.p; d.cand_list = eng->cand_list.p; d.n_list = eng->n_list.p; d.tickets = eng->tickets.p;
    d.cand_v = eng->cand_v.p; d.feas = eng->feas.p; d.inserted = eng->inserted.p; d.inserted_box = eng->inserted_box.p;
    d.check = eng->check.p; d.out_count = eng->out_count.p; d.out_off = eng->out_off.p; d.check_off = eng->check_off.p;
    d.out_xyzi = eng->out_xyzi.p; d.out_label = eng->out_label.p; d.out_check = eng->out_check.p;
    d.gcell = eng->gcell.p; d.gpts = eng->gpts.p; d.gnear = eng->gnear.p; d.gscratch = eng->gscratch.p;
    d.col_off = eng->col_off.p; d.col_idx = eng->col_idx.p; d.acell = eng->acell.p; d.apts = eng->apts.p;
    d.try_obj = eng->try_obj.p; d.chunk_cnt = eng->chunk_cnt.p;
    d.od_map_off = eng->od_map_off.p; d.od_map_dims = eng->od_map_dims.p; d.stats = eng->stats.p; d.occ_far = eng->occ_far.p; d.occ_cnt = eng->occ_cnt.p;
    *out = eng;
    return R3D_OK;
}

extern "C" int r3d_engine_destroy(r3d_engine* eng) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng) return R3D_OK;
    cudaStreamSynchronize(eng->stream);
    drain_events(eng);
    for (cudaEvent_t ev : eng->event_pool) cudaEventDestroy(ev);
    if (eng->h_words) cudaFreeHost(eng->h_words);
    if (eng->h_offsets) cudaFreeHost(eng->h_offsets);
    for (int i = 0; i < R3D_MAX_SUB; ++i) {
        if (eng->sub_stream[i]) { cudaStreamSynchronize(eng->sub_stream[i]); cudaStreamDestroy(eng->sub_stream[i]); }
        if (eng->sub_done[i]) cudaEventDestroy(eng->sub_done[i]);
    }
    if (eng->ev_armed) cudaEventDestroy(eng->ev_armed);
    if (eng->ev_wait) cudaEventDestroy(eng->ev_wait);
    for (int i = 0; i < R3D_MAX_SUB; ++i) if (eng->round_graph[i].exec) cudaGraphExecDestroy(eng->round_graph[i].exec);
    cudaStreamDestroy(eng->stream);
    delete eng;
    return R3D_OK;
}

extern "C" int r3d_engine_set_yaw_tables(r3d_engine* eng, const double* cos_k, const double* sin_k) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !cos_k || !sin_k) return r3d_fail(R3D_ERR_ARG, "r3d_engine_set_yaw_tables: null argument");
    const size_t n = (size_t)eng->dev.K + 1;
    R3D_CUDA(cudaMemcpy(eng->cos_k.p, cos_k, n * sizeof(double), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->sin_k.p, sin_k, n * sizeof(double), cudaMemcpyHostToDevice));
    eng->yaw_set = true;
    return R3D_OK;
}

extern "C" int r3d_engine_set_objects(r3d_engine* eng, const r3d_object_db* db) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !db || db->n_objects <= 0) return r3d_fail(R3D_ERR_ARG, "r3d_engine_set_objects: bad argument");
    EngineDev& d = eng->dev;
    const int n = db->n_objects;
    const int64_t total = db->point_offsets[n];
    std::vector<double> x(total), y(total), z(total);
    std::vector<float> in(total);
    std::vector<unsigned> lab(total);
    for (int64_t i = 0; i < total; ++i) {
        const double* p = db->points5 + i * 5;
        x[i] = p[0]; y[i] = p[1]; z[i] = p[2]; in[i] = (float)p[3]; lab[i] = (unsigned)p[4];
    }
    std::vector<ObjBox> ob(n);
    int max_pts = 1;
    for (int o = 0; o < n; ++o) {
        const double* bx = db->boxes + (size_t)o * 8;
        ObjBox& t = ob[o];
        t.cx = bx[0]; t.cy = bx[1]; t.cz = bx[2]; t.a = bx[3]; t.b = bx[4];
        t.length = bx[5]; t.width = bx[6]; t.height = bx[7];
        t.rho = std::hypot(t.cx, t.cy); t.psi0 = std::atan2(t.cy, t.cx);
        t.reach = 0.5 * std::hypot(t.length, t.width) + 0.05;
        t.cls = db->class_index[o];
        t.first = (int)db->point_offsets[o]; t.count = (int)(db->point_offsets[o + 1] - db->point_offsets[o]);
        t.pad = 0;
        t.eu0 = t.ev0 = t.ez0 = 1e300; t.eu1 = t.ev1 = t.ez1 = -1e300;
        for (int i = t.first; i < t.first + t.count; ++i) {
            const double dx = x[i] - t.cx, dy = y[i] - t.cy;
            const double u = dx * t.a + dy * t.b, v = -dx * t.b + dy * t.a, w = z[i] - t.cz;
            t.eu0 = std::min(t.eu0, u); t.eu1 = std::max(t.eu1, u); t.ev0 = std::min(t.ev0, v); t.ev1 = std::max(t.ev1, v);
            t.ez0 = std::min(t.ez0, w); t.ez1 = std::max(t.ez1, w);
        }
        t.eu0 -= 1e-6; t.ev0 -= 1e-6; t.ez0 -= 1e-6; t.eu1 += 1e-6; t.ev1 += 1e-6; t.ez1 += 1e-6;
        if (t.cls < 0 || t.cls >= d.n_classes) return r3d_fail(R3D_ERR_ARG, "r3d_engine_set_objects: class index out of range");
        max_pts = std::max(max_pts, t.count);
    }
    if (max_pts > 32768) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_set_objects: a cut object has more than 32768 points");
    d.max_obj_points = max_pts; d.n_objects = n;
    TRY(eng->obj_x.alloc(total)); TRY(eng->obj_y.alloc(total)); TRY(eng->obj_z.alloc(total)); TRY(eng->obj_i.alloc(total));
    TRY(eng->obj_label.alloc(total)); TRY(eng->obj.alloc(n)); TRY(eng->class_list_off.alloc(d.n_classes + 1));
    const int nl = db->class_list_offsets[d.n_classes];
    TRY(eng->class_list.alloc(std::max(nl, 1)));
    TRY(eng->unplaceable.alloc((size_t)d.B * ((n + 31) / 32)));
    TRY(eng->occ_pix.alloc((size_t)d.B * OCC_G * max_pts)); TRY(eng->sel_pix.alloc((size_t)d.B * max_pts));
    R3D_CUDA(cudaMemcpy(eng->obj_x.p, x.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->obj_y.p, y.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->obj_z.p, z.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->obj_i.p, in.data(), total * sizeof(float), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->obj_label.p, lab.data(), total * sizeof(unsigned), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->obj.p, ob.data(), n * sizeof(ObjBox), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->class_list_off.p, db->class_list_offsets, (d.n_classes + 1) * sizeof(int), cudaMemcpyHostToDevice));
    R3D_CUDA(cudaMemcpy(eng->class_list.p, db->class_list, nl * sizeof(int), cudaMemcpyHostToDevice));
    d.obj_x = eng->obj_x.p; d.obj_y = eng->obj_y.p; d.obj_z = eng->obj_z.p; d.obj_i = eng->obj_i.p; d.obj_label = eng->obj_label.p;
    d.obj = eng->obj.p; d.class_list_off = eng->class_list_off.p; d.class_list = eng->class_list.p;
    d.unplaceable = eng->unplaceable.p; d.occ_pix = eng->occ_pix.p; d.sel_pix = eng->sel_pix.p;
    const size_t sel_smem = select_smem_bytes(std::min(max_pts, SEL_SMEM_PTS));
    d.sel_key_cap = next_pow2(max_pts);
    if (max_pts > std::min(SEL_SMEM_PTS, walk_od::WALK_SEL_PTS)) {        // objects the selection cannot keep in shared memory
        TRY(eng->sel_keys.alloc((size_t)d.B * d.sel_key_cap)); TRY(eng->sel_r.alloc((size_t)d.B * max_pts));
    }
    d.sel_keys = eng->sel_keys.p; d.sel_r = eng->sel_r.p;
    R3D_CUDA(cudaFuncSetAttribute(k_select_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    R3D_CUDA(cudaFuncSetAttribute(k_occl_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)occl_smem_bytes(d)));
    if (onmap_smem_bytes(d.K) > 100 * 1024) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_set_objects: too many yaw steps for the placement kernel's shared memory");
    R3D_CUDA(cudaFuncSetAttribute(k_onmap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)onmap_smem_bytes(d.K)));
    if (grid_near_smem(d.G) > 200 * 1024) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_set_objects: road-level grid too large for shared memory");
    R3D_CUDA(cudaFuncSetAttribute(k_grid_near_bits, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_near_smem(d.G)));
    if (std::max(walk_od::walk_smem_layout(d.K, d.dwords).total, walk_ss::walk_smem_layout(d.K, d.dwords).total) > 200 * 1024)
        return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_set_objects: range image / yaw steps too large for the walker's shared memory");
    R3D_CUDA(cudaFuncSetAttribute(walk_od::k_scan_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)walk_od::walk_smem_layout(d.K, d.dwords).total));
    R3D_CUDA(cudaFuncSetAttribute(walk_ss::k_scan_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)walk_ss::walk_smem_layout(d.K, d.dwords).total));
    eng->objects_set = true;
    return R3D_OK;
}

extern "C" int r3d_engine_set_ss_map(r3d_engine* eng, const uint8_t* map, int32_t size_x, int32_t size_y, int64_t move_x,
                                     int64_t move_y) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !map || size_x <= 0 || size_y <= 0) return r3d_fail(R3D_ERR_ARG, "r3d_engine_set_ss_map: bad argument");
    TRY(eng->ss_map.alloc((size_t)size_x * size_y));
    R3D_CUDA(cudaMemcpy(eng->ss_map.p, map, (size_t)size_x * size_y, cudaMemcpyHostToDevice));
    eng->dev.ss_map = eng->ss_map.p; eng->dev.ss_sx = size_x; eng->dev.ss_sy = size_y;
    eng->dev.ss_move_x = move_x; eng->dev.ss_move_y = move_y;
    return R3D_OK;
}

static int arm_batch(r3d_engine* eng, bool ingest) {
    eng->run_active = false;                     // a new batch / re-arm abandons a run that was left unfinished
    EngineDev& d = eng->dev;
    const int n = eng->n_scans;
    cudaStream_t st = eng->stream;
    k_reset_state<<<(n + 127) / 128, 128, 0, st>>>(d, n, eng->n0_arr.p, eng->nbox0_arr.p); r3d_count_launch();
    R3D_CUDA(cudaMemsetAsync(eng->dmask.p, 0, (size_t)n * d.dwords * sizeof(unsigned), st));   // vis_px masks: cleared lazily afterwards
    const int chunks = (eng->max_n0 + CHUNK - 1) / CHUNK;
    if (ingest) {
        R3D_CUDA(cudaMemsetAsync(eng->gcell.p, 0, (size_t)n * d.G * d.G * sizeof(int), st));
        R3D_CUDA(cudaMemsetAsync(eng->col_off.p, 0, (size_t)n * (d.cols + 1) * sizeof(int), st));
        R3D_CUDA(cudaMemsetAsync(eng->acell.p, 0, (size_t)n * d.G * d.G * sizeof(int), st));
        { Launcher l(eng, KID_INGEST); k_ingest_count<<<dim3(chunks, n), STREAM_THREADS, 0, st>>>(d, n); }
        {
            Launcher l(eng, KID_GRID);
            k_bucket_scan3<<<dim3(n, 3), 1024, 0, st>>>(d, n);
        }
        { Launcher l(eng, KID_SCATTER); k_scatter_project<<<dim3(chunks, n), STREAM_THREADS, 0, st>>>(d, n); }
        {
            Launcher l(eng, KID_GRID);
            k_grid_near_bits<<<n, NEAR_THREADS, grid_near_smem(d.G), st>>>(d, n);
        }
        eng->fresh = true;                       // the first range image of every scan is already projected
    } else {
        eng->fresh = false;
        k_reset_alive<<<dim3(std::max(chunks, 1), n), 256, 0, st>>>(d, n); r3d_count_launch();
        // scene boxes: drop the boxes appended by the previous run
        R3D_CUDA(cudaMemcpyAsync(eng->boxes.p, eng->boxes0.p, eng->h_boxes.size() * sizeof(Box), cudaMemcpyDeviceToDevice, st));
        k_box_tests<<<(int)((eng->h_boxes.size() + 127) / 128), 128, 0, st>>>(eng->boxes.p, eng->box_tests.p, (int)eng->h_boxes.size());
        r3d_count_launch();
    }
    eng->ran = false;
    return r3d_check_launch("arm_batch");
}

extern "C" int r3d_engine_load_batch(r3d_engine* eng, const r3d_batch* bt) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !bt) return r3d_fail(R3D_ERR_ARG, "r3d_engine_load_batch: null argument");
    if (!eng->objects_set || !eng->yaw_set) return r3d_fail(R3D_ERR_ARG, "r3d_engine_load_batch: set objects and yaw tables first");
    EngineDev& d = eng->dev;
    const int n = bt->n_scans;
    if (n <= 0 || n > d.B) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_load_batch: n_scans exceeds max_scans");
    if (!bt->point_offsets || !bt->xyzi || (!bt->labels && !bt->labels16 && !bt->labels1) || !bt->counts || !bt->perms || bt->n_events <= 0)
        return r3d_fail(R3D_ERR_ARG, "r3d_engine_load_batch: missing arrays");
    if (d.task == 1 && (!bt->poses || !d.ss_map)) return r3d_fail(R3D_ERR_ARG, "r3d_engine_load_batch: semseg needs poses and the map");
    if (d.task == 0 && (!bt->maps || !bt->map_offsets || !bt->map_dims)) return r3d_fail(R3D_ERR_ARG, "r3d_engine_load_batch: OD needs maps");
    cudaStream_t st = eng->stream;
    std::vector<int> n0(n), nb(n);
    eng->max_n0 = 0;
    for (int s = 0; s < n; ++s) {
        const int64_t cnt = bt->point_offsets[s + 1] - bt->point_offsets[s];
        if (cnt <= 0 || cnt > d.max_points) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_load_batch: scan larger than max_points (or empty)");
        n0[s] = (int)cnt; eng->max_n0 = std::max(eng->max_n0, (int)cnt);
        nb[s] = bt->box_offsets ? bt->box_offsets[s + 1] - bt->box_offsets[s] : 0;
        if (nb[s] > d.max_boxes) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_load_batch: too many scene boxes");
    }
    eng->n_scans = n;
    // scene boxes
    eng->h_boxes.assign((size_t)n * d.max_boxes, Box());
    for (int s = 0; s < n; ++s)
        for (int j = 0; j < nb[s]; ++j) {
            const double* src = bt->boxes + ((size_t)bt->box_offsets[s] + j) * R3D_BOX_DOUBLES;
            Box& b = eng->h_boxes[(size_t)s * d.max_boxes + j];
            b.cx = src[0]; b.cy = src[1]; b.cz = src[2];
            for (int i = 0; i < 9; ++i) b.m[i] = src[3 + i];
            b.length = src[12]; b.width = src[13]; b.height = src[14]; b.reach = src[15];
        }
    eng->h_nbox0 = nb;
    R3D_CUDA(cudaMemcpyAsync(eng->boxes.p, eng->h_boxes.data(), eng->h_boxes.size() * sizeof(Box), cudaMemcpyHostToDevice, st));
    R3D_CUDA(cudaMemcpyAsync(eng->boxes0.p, eng->boxes.p, eng->h_boxes.size() * sizeof(Box), cudaMemcpyDeviceToDevice, st));
    k_box_tests<<<(int)((eng->h_boxes.size() + 127) / 128), 128, 0, st>>>(eng->boxes.p, eng->box_tests.p, (int)eng->h_boxes.size());
    r3d_count_launch();
    R3D_CUDA(cudaMemcpyAsync(eng->n0_arr.p, n0.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
    R3D_CUDA(cudaMemcpyAsync(eng->nbox0_arr.p, nb.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
    if (d.task == 0) {
        const int64_t bytes = bt->map_offsets[(size_t)n * 2];
        if ((size_t)bytes > eng->od_maps.n) TRY(eng->od_maps.alloc((size_t)bytes));
        R3D_CUDA(cudaMemcpyAsync(eng->od_maps.p, bt->maps, (size_t)bytes, cudaMemcpyHostToDevice, st));
        R3D_CUDA(cudaMemcpyAsync(eng->od_map_off.p, bt->map_offsets, ((size_t)n * 2 + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
        R3D_CUDA(cudaMemcpyAsync(eng->od_map_dims.p, bt->map_dims, (size_t)n * 8 * sizeof(int), cudaMemcpyHostToDevice, st));
        d.od_maps = eng->od_maps.p;
    } else {
        R3D_CUDA(cudaMemcpyAsync(eng->poses.p, bt->poses, (size_t)n * 16 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    for (int sc = 0; sc < n; ++sc) {                  // a schedule must fit the per-scan event / insertion records
        long long want = 0;
        for (int c = 0; c < d.n_classes; ++c) want += std::max(bt->counts[(size_t)sc * d.n_classes + c], 0);
        if (want + 1 > d.max_events)
            return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_load_batch: a scan's schedule asks for more objects than max_events - 1 "
                                              "(create the engine with max_events = objects per scan + 1)");
    }
    R3D_CUDA(cudaMemcpyAsync(eng->counts.p, bt->counts, (size_t)n * d.n_classes * sizeof(int), cudaMemcpyHostToDevice, st));
    const size_t nperm = (size_t)n * bt->n_events * d.n_classes * d.max_tries;
    if (nperm > eng->perms.n) TRY(eng->perms.alloc(nperm));
    R3D_CUDA(cudaMemcpyAsync(eng->perms.p, bt->perms, nperm * sizeof(int), cudaMemcpyHostToDevice, st));
    d.perms = eng->perms.p; d.n_perm_events = bt->n_events;
    // points: packed host rows -> per-scan strided device rows (one pitched copy when every scan has the same size)
    bool uniform = true;
    for (int s = 1; s < n; ++s) uniform &= n0[s] == n0[0];
    const bool packed = bt->labels == nullptr;         // 16-bit labels / Road bits: staged contiguously, expanded on the device
    const bool bits = packed && bt->labels16 == nullptr;
    if (bits && (d.task != 0 || bt->point_offsets[0] != 0))
        return r3d_fail(R3D_ERR_ARG, "r3d_engine_load_batch: labels1 (one Road bit per point) is for object detection batches that start at point 0");
    if (packed) {
        const size_t total = (size_t)(bt->point_offsets[n] - bt->point_offsets[0]);
        if ((size_t)n + 1 > eng->pt_off.n) TRY(eng->pt_off.alloc((size_t)d.B + 1));
        if (bits) {
            const size_t nbytes = (total + 7) / 8;
            if (nbytes > eng->label1.n) TRY(eng->label1.alloc(nbytes));
            R3D_CUDA(cudaMemcpyAsync(eng->label1.p, bt->labels1, nbytes, cudaMemcpyHostToDevice, st));
        } else {
            if (total > eng->label16.n) TRY(eng->label16.alloc(total));
            R3D_CUDA(cudaMemcpyAsync(eng->label16.p, bt->labels16 + bt->point_offsets[0], total * sizeof(unsigned short), cudaMemcpyHostToDevice, st));
        }
        R3D_CUDA(cudaMemcpyAsync(eng->pt_off.p, bt->point_offsets, ((size_t)n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    }
    if (uniform) {
        R3D_CUDA(cudaMemcpy2DAsync(eng->xyzi.p, (size_t)d.max_points * sizeof(float4), bt->xyzi + bt->point_offsets[0] * 4,
                                   (size_t)n0[0] * sizeof(float4), (size_t)n0[0] * sizeof(float4), n, cudaMemcpyHostToDevice, st));
        if (!packed)
            R3D_CUDA(cudaMemcpy2DAsync(eng->label.p, (size_t)d.P * sizeof(unsigned), bt->labels + bt->point_offsets[0],
                                       (size_t)n0[0] * sizeof(unsigned), (size_t)n0[0] * sizeof(unsigned), n, cudaMemcpyHostToDevice, st));
    } else {
        for (int s = 0; s < n; ++s) {
            const int64_t o = bt->point_offsets[s];
            R3D_CUDA(cudaMemcpyAsync(eng->xyzi.p + (size_t)s * d.max_points, bt->xyzi + o * 4, (size_t)n0[s] * sizeof(float4), cudaMemcpyHostToDevice, st));
            if (!packed)
                R3D_CUDA(cudaMemcpyAsync(eng->label.p + (size_t)s * d.P, bt->labels + o, (size_t)n0[s] * sizeof(unsigned), cudaMemcpyHostToDevice, st));
        }
    }
    if (packed) {
        const int chunks = (eng->max_n0 + CHUNK - 1) / CHUNK;
        if (bits) k_expand_label_bits<<<dim3(chunks, n), STREAM_THREADS, 0, st>>>(eng->label1.p, eng->pt_off.p, eng->label.p, d.P, n, (unsigned)d.road_label);
        else k_widen_labels<<<dim3(chunks, n), STREAM_THREADS, 0, st>>>(eng->label16.p, eng->pt_off.p, eng->label.p, d.P, n);
        r3d_count_launch();
    }
    // the std::vectors above are pageable: the copies from them completed before cudaMemcpyAsync returned
    eng->batch_loaded = true;
    return arm_batch(eng, true);
}

extern "C" int r3d_engine_reset_batch(r3d_engine* eng) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !eng->batch_loaded) return r3d_fail(R3D_ERR_ARG, "r3d_engine_reset_batch: no batch loaded");
    return arm_batch(eng, false);
}

extern "C" int r3d_engine_rearm_batch(r3d_engine* eng, int from_raw_points) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !eng->batch_loaded) return r3d_fail(R3D_ERR_ARG, "r3d_engine_rearm_batch: no batch loaded");
    if (!from_raw_points) return arm_batch(eng, false);
    // the whole device path from the resident float4 points again: scene boxes, spherical ingest, spatial indices
    R3D_CUDA(cudaMemcpyAsync(eng->boxes.p, eng->boxes0.p, eng->h_boxes.size() * sizeof(Box), cudaMemcpyDeviceToDevice, eng->stream));
    k_box_tests<<<(int)((eng->h_boxes.size() + 127) / 128), 128, 0, eng->stream>>>(eng->boxes.p, eng->box_tests.p, (int)eng->h_boxes.size());
    r3d_count_launch();
    return arm_batch(eng, true);
}

// view of the device data model that starts at scan b0: every per-scan array is [scan][...], so offsetting the base
// pointers lets the kernels run unchanged on a contiguous sub-batch
static EngineDev sub_view(const EngineDev& d, int b0) {
    if (b0 == 0) return d;
    EngineDev v = d;
    const size_t b = (size_t)b0;
    const size_t gg = (size_t)d.G * d.G, k1 = (size_t)d.K + 1, ww = (size_t)d.map_window * d.map_window / 32;
#define R3D_OFF(field, stride) if (d.field) v.field = d.field + b * (size_t)(stride)
    R3D_OFF(gcell, gg); R3D_OFF(gpts, d.max_points); R3D_OFF(gnear, gg); R3D_OFF(gscratch, gg); R3D_OFF(acell, gg); R3D_OFF(apts, d.max_points);
    R3D_OFF(xyzi, d.max_points); R3D_OFF(tail_x, d.max_inserted); R3D_OFF(tail_y, d.max_inserted);
    R3D_OFF(tail_z, d.max_inserted); R3D_OFF(tail_i, d.max_inserted); R3D_OFF(label, d.P); R3D_OFF(r, d.P); R3D_OFF(el, d.P);
    R3D_OFF(col, d.P); R3D_OFF(pix, d.P); R3D_OFF(alive, d.P); R3D_OFF(zraw, d.hw); R3D_OFF(obj_raw, d.hw);
    R3D_OFF(smooth, d.hw); R3D_OFF(dmask, d.dwords); R3D_OFF(vmask, d.dwords); R3D_OFF(st, 1); R3D_OFF(gate_update, 1);
    R3D_OFF(gate_try, 1); R3D_OFF(gate_apply, 1); R3D_OFF(gate_full, 1); R3D_OFF(gate_patch, 1); R3D_OFF(cf_rect, 4);
    R3D_OFF(work_cnt, 4); R3D_OFF(full_list, 1); R3D_OFF(cf_tasks, d.cf_tiles); R3D_OFF(need2, 1);
    R3D_OFF(col_off, d.cols + 1); R3D_OFF(col_idx, d.max_points); R3D_OFF(far_arr, 1); R3D_OFF(boxes, d.max_boxes);
    R3D_OFF(box_tests, d.max_boxes); R3D_OFF(od_map_off, 2); R3D_OFF(od_map_dims, 8); R3D_OFF(poses, 16); R3D_OFF(occ_win, ww);
    R3D_OFF(counts, d.n_classes); R3D_OFF(perms, (size_t)d.n_perm_events * d.n_classes * d.max_tries);
    R3D_OFF(unplaceable, (d.n_objects + 31) / 32); R3D_OFF(try_obj, 1); R3D_OFF(cand_flags, k1); R3D_OFF(cand_level, k1);
    R3D_OFF(cand_v, k1); R3D_OFF(cand_list, k1); R3D_OFF(n_list, 1); R3D_OFF(tickets, 4); R3D_OFF(feas, d.K);
    R3D_OFF(occ_pix, (size_t)OCC_G * d.max_obj_points); R3D_OFF(sel_pix, d.max_obj_points);
    R3D_OFF(sel_keys, d.sel_key_cap); R3D_OFF(sel_r, d.max_obj_points); R3D_OFF(inserted, (size_t)d.max_events * 4);
    R3D_OFF(inserted_box, (size_t)d.max_events * 8); R3D_OFF(check, (size_t)d.max_inserted * 5); R3D_OFF(chunk_cnt, d.max_chunks);
    R3D_OFF(out_count, 1); R3D_OFF(occ_far, OCC_FAR_CAP + 1); R3D_OFF(occ_cnt, (size_t)d.map_window * d.map_window);
#undef R3D_OFF
    return v;
}

extern "C" int r3d_engine_run_until(r3d_engine* eng, int stop_at, int* still_running);

extern "C" int r3d_engine_run(r3d_engine* eng) { return r3d_engine_run_until(eng, 0, nullptr); }

// The default execution model: batch-wide streaming kernels for the first range image of every scan, then ONE launch
// of the per-scan walker (r3d_k_walk.cuh) that takes every scan through all its slots and tries, then the output
// compaction.  Everything is queued on the engine stream; the host neither polls nor synchronises.
static int run_walker(r3d_engine* eng) {
    const EngineDev& d = eng->dev;
    const int n = eng->n_scans;
    cudaStream_t st = eng->stream;
    const int chunks0 = (eng->max_n0 + CHUNK - 1) / CHUNK;
    const int chunks_all = (eng->max_n0 + d.max_inserted + CHUNK - 1) / CHUNK;
#ifndef R3D_FULL_Y
#define R3D_FULL_Y 32
#endif
    const int full_y = std::min(n, R3D_FULL_Y);
    const bool fresh = eng->fresh;
    eng->fresh = false;
    { Launcher l(eng, KID_PREP); walk_od::k_walk_prepare<<<std::min((n * d.cf_tiles + 255) / 256, eng->n_sms * 4), 256, 0, st>>>(d, n, fresh ? 0 : 1); }
    if (!fresh) {                                // re-armed without ingest: rebuild the first range image from the caches
        { Launcher l(eng, KID_MINMAX0); k_minmax<<<dim3(chunks0, full_y), STREAM_THREADS, 0, st>>>(d, n); }
        { Launcher l(eng, KID_CLEAR0); k_clear_images<<<dim3(32, full_y), STREAM_THREADS, 0, st>>>(d, n); }
        { Launcher l(eng, KID_PROJECT0); k_project<<<dim3(chunks0, full_y), STREAM_THREADS, 0, st>>>(d, n); }
    }
    {
        Launcher l(eng, KID_CLOSEFILL0);
        const int cf_grid = std::min(n * d.cf_tiles, eng->n_sms * 4);
        if (eng->zraw_tmap_ok) {                 // tile loads by the TMA unit
            k_close_fill_tma<<<cf_grid, CF_THREADS, 0, st>>>(eng->zraw_tmap, d.rows, d.cols, (int64_t)d.hw, d.smooth, d.far_arr, d.cf_tasks,
                                                              d.work_cnt + 1);
        } else if ((d.cols & 1) == 0 && ((uintptr_t)d.zraw & 15) == 0) {
            k_close_fill_raw_pipelined<<<cf_grid, CF_THREADS, 0, st>>>(d.zraw, d.rows, d.cols, (int64_t)d.hw, d.smooth, d.far_arr, d.cf_tasks,
                                                                        d.work_cnt + 1);
        } else {
            RawImage in{d.zraw};
            k_close_fill_tasks<RawImage><<<cf_grid, CF_THREADS, 0, st>>>(in, d.rows, d.cols, (int64_t)d.hw, d.smooth, d.far_arr, d.cf_tasks,
                                                                         d.work_cnt + 1);
        }
    }
    if (d.task == 1) {
        R3D_CUDA(cudaMemsetAsync(eng->occ_win.p, 0, (size_t)n * ((size_t)d.map_window * d.map_window / 32) * sizeof(unsigned), st));
        R3D_CUDA(cudaMemsetAsync(eng->occ_cnt.p, 0, (size_t)n * (size_t)d.map_window * d.map_window * sizeof(unsigned), st));
        Launcher l(eng, KID_ADJUST); k_adjust_map<<<dim3(chunks0, n), STREAM_THREADS, 0, st>>>(d, n, 1);
    }
    {
        Launcher l(eng, KID_WALK);               // CTA shape per task: see the head of r3d_k_walk.cuh
        if (d.task == 1) walk_ss::k_scan_walk<<<n, walk_ss::WALK_THREADS, walk_ss::walk_smem_layout(d.K, d.dwords).total, st>>>(d, n);
        else walk_od::k_scan_walk<<<n, walk_od::WALK_THREADS, walk_od::walk_smem_layout(d.K, d.dwords).total, st>>>(d, n);
    }
    {
        Launcher l(eng, KID_OUT);
        k_out_count<<<dim3(chunks_all, n), STREAM_THREADS, 0, st>>>(d, n);
        k_out_offsets<<<1, 1024, 0, st>>>(d, n, chunks_all);
        k_out_write<<<dim3(chunks_all, n), STREAM_THREADS, 0, st>>>(d, n);
        r3d_count_launch(2);
    }
    R3D_CUDA(cudaMemcpyAsync(eng->h_offsets, eng->out_off.p, (n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    R3D_CUDA(cudaMemcpyAsync(eng->h_offsets + (d.B + 1), eng->check_off.p, (n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    R3D_CUDA(cudaMemcpyAsync(eng->h_offsets + 2 * (d.B + 1), eng->stats.p + 8, sizeof(long long), cudaMemcpyDeviceToHost, st));
    eng->ran = true; eng->walked = true;
    return r3d_check_launch("r3d_engine_run");
}

extern "C" int r3d_engine_run_until(r3d_engine* eng, int stop_at, int* still_running) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !eng->batch_loaded) return r3d_fail(R3D_ERR_ARG, "r3d_engine_run: no batch loaded");
    if (still_running) *still_running = 0;
    if (!eng->staged) return run_walker(eng);
    eng->walked = false;
    const EngineDev& d0 = eng->dev;
    const int n = eng->n_scans;
    cudaStream_t st = eng->stream;
    const int P_live = eng->max_n0 + d0.max_inserted;
    const int chunks_all = (P_live + CHUNK - 1) / CHUNK;
    const int sel_pts = std::min(d0.max_obj_points, SEL_SMEM_PTS);
    const size_t sel_smem = select_smem_bytes(sel_pts);
    const int key_cap = next_pow2(sel_pts);
    const size_t onmap_smem = onmap_smem_bytes(d0.K);
    const int max_rounds = d0.max_events * (3 * d0.max_tries + 2) + 8;
    // The batch is advanced as n_sub contiguous sub-batches, each on its own stream: every kernel of a round is a
    // short chain of dependent loads that leaves most of the SMs' issue slots idle, so the rounds of different
    // sub-batches overlap on the device.
    typedef r3d_engine::Sub Sub;
    Sub* sub = eng->sub;
    if (!eng->run_active) {
        eng->run_nsub = std::max(1, std::min(eng->n_sub, n));
        eng->run_done = 0; eng->run_rounds = 0;
        R3D_CUDA(cudaEventRecord(eng->ev_armed, st));
        for (int i = 0; i < eng->run_nsub; ++i) {
            Sub& s = sub[i];
            s.b0 = (int)((long long)n * i / eng->run_nsub); s.n = (int)((long long)n * (i + 1) / eng->run_nsub) - s.b0;
            s.round = 0; s.done = false; s.left = s.n; s.d = sub_view(d0, s.b0); s.st = eng->sub_stream[i];
            s.d.active_count = eng->active_count.p + (size_t)i * 128;
            s.d.host_word = eng->d_words + (size_t)i * 64;
            s.d.round_ctl = eng->round_ctl.p + (size_t)i * 32;
            if (eng->seq > 0xF0000000u) eng->seq = 0;                       // sequence numbers are never 0 (= a fresh word)
            s.seq_base = eng->seq + 1; eng->seq += (unsigned)max_rounds + 8u;
            R3D_CUDA(cudaStreamWaitEvent(s.st, eng->ev_armed, 0));
            R3D_CUDA(cudaMemsetAsync(s.d.active_count, 0, 128 * sizeof(int), s.st));
            k_set_round<<<1, 1, 0, s.st>>>(s.d.round_ctl, s.seq_base); r3d_count_launch();
        }
        eng->run_active = true;
    }
    const int nsub = eng->run_nsub;
    int& n_done = eng->run_done;
    int& rounds_max = eng->run_rounds;
    const int task_ctas = std::max(eng->n_sms, eng->n_sms * TASK_CTAS_PER_SM / nsub);
    // The k_ctrl of every round publishes (sequence number, unfinished scans) into mapped host memory.  The host keeps
    // up to `ahead` rounds of a sub-batch in flight: round r is launched once the word of round r - ahead is there
    // (rounds launched after the last scan finished are gated off on the device and cost a few microseconds each),
    // so neither the device nor the other sub-batches ever wait for a word that is stuck behind bulk PCIe traffic.
    const int ahead = 3;
    auto poll_round = [&](int i, int round, long long& left) -> int {       // 1: published, 0: not yet
        volatile unsigned long long* w = eng->h_words + (size_t)i * 64 + (round & 63);
        const unsigned long long v = *w;
        if ((unsigned)(v >> 32) == sub[i].seq_base + (unsigned)round) { left = (long long)(v & 0xffffffffull); return 1; }
        return 0;
    };
    // the launch sequence of one round; also what a round graph is captured from
    auto launch_round = [&](const EngineDev& d, int ns, cudaStream_t ss, bool r0) {
        const size_t pref_smem = (size_t)(ns + 1) * sizeof(int);
        // the three full re-projection kernels and close/fill walk the work lists k_update wrote (all scans in round
        // 0, a handful later): grids sized for "some scans", not for every (chunk, scan) / (tile, scan) pair
#ifndef R3D_FULL_Y
#define R3D_FULL_Y 32
#endif
        const int full_y = std::min(ns, R3D_FULL_Y);
        const int cf_grid = std::min(ns * d.cf_tiles, eng->n_sms * 8);
        { Launcher l(eng, KID_CTRL, ss); k_ctrl<<<ns, 32, 0, ss>>>(d, ns); }
        { Launcher l(eng, KID_UPDATE, ss); k_update<<<dim3(UPDATE_G, ns), UPDATE_THREADS, 0, ss>>>(d, ns); }
        { Launcher l(eng, r0 ? KID_MINMAX0 : KID_MINMAX, ss); k_minmax<<<dim3(chunks_all, full_y), STREAM_THREADS, 0, ss>>>(d, ns); }
        { Launcher l(eng, r0 ? KID_CLEAR0 : KID_CLEAR, ss); k_clear_images<<<dim3(32, full_y), STREAM_THREADS, 0, ss>>>(d, ns); }
        { Launcher l(eng, r0 ? KID_PROJECT0 : KID_PROJECT, ss); k_project<<<dim3(chunks_all, full_y), STREAM_THREADS, 0, ss>>>(d, ns); }
        {
            Launcher l(eng, r0 ? KID_CLOSEFILL0 : KID_CLOSEFILL, ss);
            if ((d.cols & 1) == 0 && ((uintptr_t)d.zraw & 15) == 0) {
                k_close_fill_raw_pipelined<<<std::min(cf_grid, eng->n_sms * 4), CF_THREADS, 0, ss>>>(d.zraw, d.rows, d.cols, (int64_t)d.hw, d.smooth, d.far_arr,
                                                                           d.cf_tasks, d.work_cnt + 1);
            } else {
                RawImage in{d.zraw};
                k_close_fill_tasks<RawImage><<<cf_grid, CF_THREADS, 0, ss>>>(in, d.rows, d.cols, (int64_t)d.hw, d.smooth, d.far_arr,
                                                                             d.cf_tasks, d.work_cnt + 1);
            }
        }
        if (d.task == 1) { Launcher l(eng, KID_ADJUST, ss); k_adjust_map<<<dim3(chunks_all, ns), STREAM_THREADS, 0, ss>>>(d, ns); }
        { Launcher l(eng, KID_ONMAP, ss); k_onmap<<<ns, TRY_THREADS, onmap_smem, ss>>>(d, ns); }
        const int ph = d.cand_window > 0 ? 1 : 0;           // OD: first window of candidates, then the rest where needed
        if (d.task == 0) { Launcher l(eng, KID_ONMAP, ss); k_onmap_full<<<task_ctas, TASK_THREADS, pref_smem, ss>>>(d, ns, ph); }
        { Launcher l(eng, KID_HEIGHT, ss); k_road_level<<<task_ctas, TASK_THREADS, pref_smem, ss>>>(d, ns, ph); }
        if (d.task == 1) {
            Launcher l(eng, KID_ONMAP, ss);
            if (d.K <= SS_MAX_K) k_onmap_ss<<<ns, 1024, 0, ss>>>(d, ns); else k_onmap_ss_seq<<<ns, 1024, 0, ss>>>(d, ns);
        }
        { Launcher l(eng, KID_COLLIDE, ss); k_collide<<<task_ctas, TASK_THREADS, pref_smem, ss>>>(d, ns, ph); }
        { Launcher l(eng, KID_OCCL, ss); k_occl_count<<<dim3(OCC_G, ns), 128, occl_smem_bytes(d), ss>>>(d, ns, ph); }
        if (ph) {
            { Launcher l(eng, KID_CTRL, ss); k_phase_gate<<<(ns + 127) / 128, 128, 0, ss>>>(d, ns); }
            { Launcher l(eng, KID_ONMAP, ss); k_onmap_full<<<task_ctas, TASK_THREADS, pref_smem, ss>>>(d, ns, 2); }
            { Launcher l(eng, KID_HEIGHT, ss); k_road_level<<<task_ctas, TASK_THREADS, pref_smem, ss>>>(d, ns, 2); }
            { Launcher l(eng, KID_COLLIDE, ss); k_collide<<<task_ctas, TASK_THREADS, pref_smem, ss>>>(d, ns, 2); }
            { Launcher l(eng, KID_OCCL, ss); k_occl_count<<<dim3(OCC_G, ns), 128, occl_smem_bytes(d), ss>>>(d, ns, 2); }
        }
        { Launcher l(eng, KID_SELECT, ss); k_select_emit<<<ns, 512, sel_smem, ss>>>(d, ns, key_cap, sel_pts); }
    };
    const bool graphs = eng->use_graphs && !eng->profile;
    if (graphs) {
        for (int i = 0; i < nsub; ++i) {
            r3d_engine::RoundGraph& g = eng->round_graph[i];
            const Sub& s = sub[i];
            if (g.exec && g.ns == s.n && g.chunks_all == chunks_all && g.task_ctas == task_ctas && g.sel_pts == sel_pts &&
                memcmp(&g.d, &s.d, sizeof(EngineDev)) == 0) continue;
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            cudaGraph_t graph = nullptr;
            R3D_CUDA(cudaStreamBeginCapture(s.st, cudaStreamCaptureModeThreadLocal));
            eng->capturing = true;
            launch_round(s.d, s.n, s.st, false);
            eng->capturing = false;
            R3D_CUDA(cudaStreamEndCapture(s.st, &graph));
            size_t n_nodes = 0;
            R3D_CUDA(cudaGraphGetNodes(graph, nullptr, &n_nodes));
            const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) { g.exec = nullptr; return r3d_fail_cuda(ie, "r3d_engine_run: cudaGraphInstantiate"); }
            memcpy(&g.d, &s.d, sizeof(EngineDev));
            g.ns = s.n; g.chunks_all = chunks_all; g.task_ctas = task_ctas; g.sel_pts = sel_pts; g.kernels = (int)n_nodes;
        }
    }
    unsigned idle_polls = 0;
    while (n_done < nsub) {
        bool progressed = false;
        for (int i = 0; i < nsub; ++i) {
            Sub& s = sub[i];
            if (s.done) continue;
            if (s.round >= ahead) {                   // the word of round (s.round - ahead) gates the next launch
                long long left = 0;
                const int got = poll_round(i, s.round - ahead, left);
                if (got == 0) continue;
                s.left = left;
                if (left == 0) { s.done = true; ++n_done; progressed = true; rounds_max = std::max(rounds_max, s.round - ahead + 1); continue; }
            }
            progressed = true;
            if (s.round >= max_rounds) { eng->run_active = false; return r3d_fail(R3D_ERR_ARG, "r3d_engine_run: round limit reached"); }
            if (graphs) {
                R3D_CUDA(cudaGraphLaunch(eng->round_graph[i].exec, s.st));
                r3d_count_launch(eng->round_graph[i].kernels);
            } else {
                launch_round(s.d, s.n, s.st, s.round == 0);
            }
            s.round += 1;
        }
        if (stop_at > 0 && n_done < nsub) {             // hand the device to the next engine once only a tail is left
            long long left_total = 0;
            for (int i = 0; i < nsub; ++i) left_total += sub[i].done ? 0 : sub[i].left;
            if (left_total <= stop_at) { if (still_running) *still_running = 1; return R3D_OK; }
        }
        if (progressed) { idle_polls = 0; continue; }
        // nothing to launch: every unfinished sub-batch waits for a word.  Poll briefly, then sleep in short steps
        // instead of burning a host core per engine thread (8 ranks x 3 pipelined engines share the box's cores)
        if (++idle_polls > 64) std::this_thread::sleep_for(std::chrono::microseconds(20));
        if ((idle_polls & 0x3ff) == 0x3ff) {
            for (int i = 0; i < nsub; ++i) {
                if (sub[i].done) continue;
                const cudaError_t q = cudaStreamQuery(sub[i].st);
                long long left = 0;
                if ((q != cudaSuccess && q != cudaErrorNotReady) ||
                    (q == cudaSuccess && poll_round(i, sub[i].round - ahead, left) == 0))        // drained without the word
                    { eng->run_active = false; return r3d_fail_cuda(q == cudaSuccess ? cudaErrorUnknown : q, "r3d_engine_run: device error while waiting for a round"); }
            }
        }
    }
    eng->last_rounds = rounds_max;
    eng->run_active = false;
    for (int i = 0; i < nsub; ++i) {
        R3D_CUDA(cudaEventRecord(eng->sub_done[i], sub[i].st));
        R3D_CUDA(cudaStreamWaitEvent(st, eng->sub_done[i], 0));
    }
    {
        const EngineDev& d = d0;
        Launcher l(eng, KID_OUT);
        k_out_count<<<dim3(chunks_all, n), STREAM_THREADS, 0, st>>>(d, n);
        k_out_offsets<<<1, 1024, 0, st>>>(d, n, chunks_all);
        k_out_write<<<dim3(chunks_all, n), STREAM_THREADS, 0, st>>>(d, n);
        r3d_count_launch(2);
    }
    R3D_CUDA(cudaMemcpyAsync(eng->h_offsets, eng->out_off.p, (n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    R3D_CUDA(cudaMemcpyAsync(eng->h_offsets + (d0.B + 1), eng->check_off.p, (n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    eng->ran = true;
    return r3d_check_launch("r3d_engine_run");
}

extern "C" int r3d_engine_set_sub_batches(r3d_engine* eng, int n_sub) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || n_sub < 1 || n_sub > R3D_MAX_SUB) return r3d_fail(R3D_ERR_ARG, "r3d_engine_set_sub_batches: 1..16");
    eng->n_sub = n_sub;
    return R3D_OK;
}

extern "C" int r3d_engine_sync(r3d_engine* eng) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng) return r3d_fail(R3D_ERR_ARG, "r3d_engine_sync: null engine");
    R3D_CUDA(engine_wait(eng));
    drain_events(eng);
    return R3D_OK;
}

extern "C" int r3d_engine_output_rows(r3d_engine* eng, int64_t* total_points, int64_t* total_check) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !eng->ran) return r3d_fail(R3D_ERR_ARG, "r3d_engine_output_rows: run first");
    R3D_CUDA(engine_wait(eng));
    if (total_points) *total_points = eng->h_offsets[eng->n_scans];
    if (total_check) *total_check = eng->h_offsets[eng->dev.B + 1 + eng->n_scans];
    return R3D_OK;
}

extern "C" int r3d_engine_output_device(r3d_engine* eng, const float** out_xyzi, const uint32_t** out_labels, const float** check,
                                        int64_t* out_offsets_host, int64_t* check_offsets_host) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !eng->ran) return r3d_fail(R3D_ERR_ARG, "r3d_engine_output_device: run first");
    R3D_CUDA(engine_wait(eng));
    const int n = eng->n_scans;
    if (out_xyzi) *out_xyzi = reinterpret_cast<const float*>(eng->out_xyzi.p);
    if (out_labels) *out_labels = eng->out_label.p;
    if (check) *check = eng->out_check.p;
    if (out_offsets_host) memcpy(out_offsets_host, eng->h_offsets, (n + 1) * sizeof(long long));
    if (check_offsets_host) memcpy(check_offsets_host, eng->h_offsets + eng->dev.B + 1, (n + 1) * sizeof(long long));
    return R3D_OK;
}

extern "C" int r3d_engine_fetch(r3d_engine* eng, r3d_batch_result* res) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !res || !eng->ran) return r3d_fail(R3D_ERR_ARG, "r3d_engine_fetch: run first");
    EngineDev& d = eng->dev;
    const int n = eng->n_scans;
    cudaStream_t st = eng->stream;
    R3D_CUDA(engine_wait(eng));
    const long long total = eng->h_offsets[n], total_check = eng->h_offsets[d.B + 1 + n];
    if (total > res->capacity_points || total_check > res->capacity_check)
        return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_fetch: result buffers too small");
    if (res->out_offsets) memcpy(res->out_offsets, eng->h_offsets, (n + 1) * sizeof(long long));
    if (res->check_offsets) memcpy(res->check_offsets, eng->h_offsets + d.B + 1, (n + 1) * sizeof(long long));
    if (res->out_xyzi && total) R3D_CUDA(cudaMemcpyAsync(res->out_xyzi, eng->out_xyzi.p, (size_t)total * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (res->out_labels && total) R3D_CUDA(cudaMemcpyAsync(res->out_labels, eng->out_label.p, (size_t)total * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    if (res->out_labels16 && total) {                  // 16-bit labels over PCIe (widened again by the caller)
        if ((size_t)total > eng->out_label16.n) TRY(eng->out_label16.alloc((size_t)d.B * d.P));
        k_narrow_labels<<<eng->n_sms * 4, 256, 0, st>>>(eng->out_label.p, eng->out_label16.p, total); r3d_count_launch();
        R3D_CUDA(cudaMemcpyAsync(res->out_labels16, eng->out_label16.p, (size_t)total * sizeof(unsigned short), cudaMemcpyDeviceToHost, st));
    }
    if (res->check_xyzil && total_check) R3D_CUDA(cudaMemcpyAsync(res->check_xyzil, eng->out_check.p, (size_t)total_check * 5 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (res->inserted) R3D_CUDA(cudaMemcpyAsync(res->inserted, eng->inserted.p, (size_t)n * d.max_events * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (res->inserted_box) R3D_CUDA(cudaMemcpyAsync(res->inserted_box, eng->inserted_box.p, (size_t)n * d.max_events * 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    std::vector<ScanState> hs(n);
    R3D_CUDA(cudaMemcpyAsync(hs.data(), eng->st.p, n * sizeof(ScanState), cudaMemcpyDeviceToHost, st));
    R3D_CUDA(engine_wait(eng));
    for (int s = 0; s < n; ++s) {
        if (res->n_inserted) res->n_inserted[s] = hs[s].n_inserted;
        if (res->status) res->status[s] = hs[s].status;
    }
    if (res->rounds) *res->rounds = eng->walked ? (int)eng->h_offsets[2 * (d.B + 1)] : eng->last_rounds;
    return R3D_OK;
}

extern "C" int r3d_engine_profile_enable(r3d_engine* eng, int on) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng) return r3d_fail(R3D_ERR_ARG, "r3d_engine_profile_enable: null engine");
    cudaStreamSynchronize(eng->stream);
    drain_events(eng);
    eng->profile = on != 0;
    for (int i = 0; i < KID_COUNT; ++i) { eng->prof_ms[i] = 0; eng->prof_launches[i] = 0; }
    R3D_CUDA(cudaMemset(eng->stats.p, 0, R3D_N_STATS * sizeof(unsigned long long)));
    return R3D_OK;
}

extern "C" int r3d_engine_stats_ex(r3d_engine* eng, uint64_t* out, int n) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !out || n < 0 || n > R3D_N_STATS) return r3d_fail(R3D_ERR_ARG, "r3d_engine_stats_ex: bad argument");
    R3D_CUDA(cudaStreamSynchronize(eng->stream));
    R3D_CUDA(cudaMemcpy(out, eng->stats.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return R3D_OK;
}

extern "C" int r3d_engine_walk_profile(r3d_engine* eng, int64_t* cycles, int32_t* tries, int n) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !cycles || !tries || n < 0 || n > eng->dev.B) return r3d_fail(R3D_ERR_ARG, "r3d_engine_walk_profile: bad argument");
    R3D_CUDA(cudaStreamSynchronize(eng->stream));
    std::vector<ScanState> st((size_t)n);
    R3D_CUDA(cudaMemcpy(st.data(), eng->dev.st, (size_t)n * sizeof(ScanState), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) { cycles[i] = st[i].walk_cycles; tries[i] = st[i].walk_tries; }
    return R3D_OK;
}

extern "C" int r3d_engine_stats(r3d_engine* eng, uint64_t* out8) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !out8) return r3d_fail(R3D_ERR_ARG, "r3d_engine_stats: null argument");
    R3D_CUDA(cudaStreamSynchronize(eng->stream));
    R3D_CUDA(cudaMemcpy(out8, eng->stats.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return R3D_OK;
}

extern "C" void* r3d_engine_stream(r3d_engine* eng) {
    if (eng) cudaSetDevice(eng->device); return eng ? (void*)eng->stream : nullptr; }

extern "C" int r3d_engine_profile_read(r3d_engine* eng, char* names_out, int names_cap, double* ms_out,
                                       int64_t* launches_out, int max_kernels, int* n_kernels_out) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng) return r3d_fail(R3D_ERR_ARG, "r3d_engine_profile_read: null engine");
    cudaStreamSynchronize(eng->stream);
    drain_events(eng);
    std::string names;
    const int n = std::min<int>(KID_COUNT, max_kernels);
    for (int i = 0; i < n; ++i) {
        names += kKernelNames[i]; names += '\n';
        if (ms_out) ms_out[i] = eng->prof_ms[i];
        if (launches_out) launches_out[i] = eng->prof_launches[i];
    }
    if (names_out && names_cap > 0) { strncpy(names_out, names.c_str(), names_cap - 1); names_out[names_cap - 1] = 0; }
    if (n_kernels_out) *n_kernels_out = n;
    return R3D_OK;
}

extern "C" int r3d_engine_debug_image(r3d_engine* eng, int scan, double* smooth_out) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || scan < 0 || scan >= eng->n_scans || !smooth_out) return r3d_fail(R3D_ERR_ARG, "r3d_engine_debug_image: bad argument");
    R3D_CUDA(cudaStreamSynchronize(eng->stream));
    R3D_CUDA(cudaMemcpy(smooth_out, eng->smooth.p + (size_t)scan * eng->dev.hw, (size_t)eng->dev.hw * sizeof(double), cudaMemcpyDeviceToHost));
    return R3D_OK;
}

extern "C" int r3d_engine_debug_candidates(r3d_engine* eng, int scan, uint8_t* flags_out, double* level_out, int32_t* visible_out) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || scan < 0 || scan >= eng->n_scans) return r3d_fail(R3D_ERR_ARG, "r3d_engine_debug_candidates: bad argument");
    const size_t K1 = eng->dev.K + 1;
    R3D_CUDA(cudaStreamSynchronize(eng->stream));
    if (flags_out) R3D_CUDA(cudaMemcpy(flags_out, eng->cand_flags.p + scan * K1, K1, cudaMemcpyDeviceToHost));
    if (level_out) R3D_CUDA(cudaMemcpy(level_out, eng->cand_level.p + scan * K1, K1 * sizeof(double), cudaMemcpyDeviceToHost));
    if (visible_out) R3D_CUDA(cudaMemcpy(visible_out, eng->cand_v.p + scan * K1, K1 * sizeof(int), cudaMemcpyDeviceToHost));
    return R3D_OK;
}

// ------------------------------------------------------------------------------- find_possible_places probe
namespace {

// the probed scan: originals stay resident for the road level but are not part of the "current scene"; the caller's
// rows become the (fp64) tail; one try is armed for `obj`
__global__ void k_probe_setup(EngineDev e, int n_scans, int scan, int obj, int n_rows) {
    const int b = blockIdx.x;
    if (b >= n_scans) return;
    ScanState& s = e.st[b];
    const bool me = b == scan;
    if (threadIdx.x == 0) {
        s.try_active = me; s.need_project = 0; s.apply_flag = 0; s.dirty = 0;
        e.gate_try[b] = me; e.gate_update[b] = me; e.gate_apply[b] = 0; e.gate_full[b] = 0; e.gate_patch[b] = 0;
        for (int i = 0; i < 4; ++i) e.tickets[(size_t)b * 4 + i] = 0u;
        if (me) {
            s.n_tail = n_rows; s.cur_obj = obj; s.cur_class = e.obj[obj].cls; s.n_feasible = 0; s.found_rank = INT_MAX;
            // the whole probed scene is one "inserted object" with an all-covering, zero-size box, so that k_collide
            // visits every tail point and no object point can ever be inside that box
            Box dummy;
            dummy.cx = dummy.cy = dummy.cz = 0.0;
            for (int i = 0; i < 9; ++i) dummy.m[i] = (i % 4 == 0) ? 1.0 : 0.0;
            dummy.length = dummy.width = dummy.height = 0.0; dummy.reach = 1e9;
            e.boxes[(size_t)b * e.max_boxes + s.n_boxes] = dummy;
            e.box_tests[(size_t)b * e.max_boxes + s.n_boxes] = make_box_test(dummy);
            e.inserted[((size_t)b * e.max_events + 0) * 4 + 3] = n_rows;
            s.n_boxes += 1; s.n_inserted = 1;
        }
    }
    if (!me) return;
    const int ww = e.map_window * e.map_window / 32;
    if (e.task == 1) for (int i = threadIdx.x; i < ww; i += blockDim.x) e.occ_win[(size_t)b * ww + i] = 0u;
    if (e.task == 1) occ_far_clear(e, b, threadIdx.x);
}

__global__ void k_probe_points(EngineDev e, int scan, int obj, int n_feasible, double* xyz_out, double* box_out) {
    const ObjBox ob = e.obj[obj];
    const size_t cb = (size_t)scan * (e.K + 1);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= e.K; k += gridDim.x * blockDim.x) {
        const YawBox yb = make_yaw_box(ob.cx, ob.cy, ob.a, ob.b, e.cos_k[k], e.sin_k[k]);
        double* bo = box_out + (size_t)k * 5;
        bo[0] = yb.cx; bo[1] = yb.cy; bo[2] = e.cand_level[cb + k]; bo[3] = yb.m00; bo[4] = yb.m10;
    }
    if (!xyz_out) return;
    const long long total = (long long)n_feasible * ob.count;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / ob.count), i = (int)(t % ob.count);
        const int k = e.feas[(size_t)scan * e.K + f];
        const double c = e.cos_k[k], sn = e.sin_k[k];
        const double x0 = e.obj_x[ob.first + i], y0 = e.obj_y[ob.first + i];
        double* o = xyz_out + t * 3;
        o[0] = sub(mul(c, x0), mul(sn, y0));
        o[1] = add(mul(sn, x0), mul(c, y0));
        o[2] = add(e.obj_z[ob.first + i], sub(e.cand_level[cb + k], ob.cz));
    }
}

}  // namespace

extern "C" int r3d_engine_probe_places(r3d_engine* eng, int scan, int object_id, const double* scene_rows9, int64_t n_rows,
                                       uint8_t* flags_out, double* box_out, double* xyz_out, int xyz_capacity,
                                       int32_t* n_feasible_out) {
    if (eng) cudaSetDevice(eng->device);
    if (!eng || !eng->batch_loaded || scan < 0 || scan >= eng->n_scans || object_id < 0 || object_id >= eng->dev.n_objects ||
        n_rows < 0 || (n_rows > 0 && !scene_rows9) || !flags_out || !box_out || !n_feasible_out)
        return r3d_fail(R3D_ERR_ARG, "r3d_engine_probe_places: bad argument");
    EngineDev d = eng->dev;
    if (n_rows > d.max_inserted) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_probe_places: scene larger than max_inserted");
    if (eng->h_nbox0[scan] + 1 > d.max_boxes) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_probe_places: max_boxes too small");
    const int n = eng->n_scans;
    cudaStream_t st = eng->stream;
    TRY(arm_batch(eng, false));
    // current scene -> fp64 tail of the probed scan; the original rows are not part of it
    std::vector<double> x(n_rows), y(n_rows), z(n_rows);
    std::vector<float> in(n_rows);
    std::vector<unsigned> lab(n_rows);
    for (int64_t i = 0; i < n_rows; ++i) {
        const double* p = scene_rows9 + i * 9;
        x[i] = p[0]; y[i] = p[1]; z[i] = p[2]; in[i] = (float)p[6]; lab[i] = (unsigned)(long long)p[7];
    }
    R3D_CUDA(cudaStreamSynchronize(st));
    std::vector<int> n0v(n);
    R3D_CUDA(cudaMemcpy(n0v.data(), eng->n0_arr.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    const int n0 = n0v[scan];
    const size_t tb = (size_t)scan * d.max_inserted, pb = (size_t)scan * d.P;
    R3D_CUDA(cudaMemsetAsync(eng->alive.p + pb, 0, n0, st));
    if (n_rows) {
        R3D_CUDA(cudaMemcpyAsync(eng->tail_x.p + tb, x.data(), n_rows * sizeof(double), cudaMemcpyHostToDevice, st));
        R3D_CUDA(cudaMemcpyAsync(eng->tail_y.p + tb, y.data(), n_rows * sizeof(double), cudaMemcpyHostToDevice, st));
        R3D_CUDA(cudaMemcpyAsync(eng->tail_z.p + tb, z.data(), n_rows * sizeof(double), cudaMemcpyHostToDevice, st));
        R3D_CUDA(cudaMemcpyAsync(eng->tail_i.p + tb, in.data(), n_rows * sizeof(float), cudaMemcpyHostToDevice, st));
        R3D_CUDA(cudaMemcpyAsync(eng->label.p + pb + n0, lab.data(), n_rows * sizeof(unsigned), cudaMemcpyHostToDevice, st));
        R3D_CUDA(cudaMemsetAsync(eng->alive.p + pb + n0, 1, n_rows, st));
    }
    k_probe_setup<<<n, 128, 0, st>>>(d, n, scan, object_id, (int)n_rows); r3d_count_launch();
    const int chunks_all = (n0 + (int)n_rows + CHUNK - 1) / CHUNK;
    if (d.task == 1) { k_adjust_map<<<dim3(chunks_all, n), STREAM_THREADS, 0, st>>>(d, n); r3d_count_launch(); }
    const int task_ctas = eng->n_sms * TASK_CTAS_PER_SM;
    const size_t pref_smem = (size_t)(n + 1) * sizeof(int);
    k_onmap<<<n, TRY_THREADS, onmap_smem_bytes(d.K), st>>>(d, n); r3d_count_launch();
    if (d.task == 0) { k_onmap_full<<<task_ctas, TASK_THREADS, pref_smem, st>>>(d, n, 0); r3d_count_launch(); }
    k_road_level<<<task_ctas, TASK_THREADS, pref_smem, st>>>(d, n, 0); r3d_count_launch();
    if (d.task == 1) { if (d.K <= SS_MAX_K) k_onmap_ss<<<n, 1024, 0, st>>>(d, n); else k_onmap_ss_seq<<<n, 1024, 0, st>>>(d, n); r3d_count_launch(); }
    k_collide<<<task_ctas, TASK_THREADS, pref_smem, st>>>(d, n, 0); r3d_count_launch();
    k_feasible_list<<<n, 128, 0, st>>>(d, n); r3d_count_launch();
    std::vector<ScanState> hs(1);
    R3D_CUDA(cudaMemcpyAsync(hs.data(), eng->st.p + scan, sizeof(ScanState), cudaMemcpyDeviceToHost, st));
    R3D_CUDA(cudaStreamSynchronize(st));
    const int nf = hs[0].n_feasible;
    *n_feasible_out = nf;
    if (hs[0].status != 0) return r3d_fail(hs[0].status, "r3d_engine_probe_places: scan status");
    if (xyz_out && nf > xyz_capacity) return r3d_fail(R3D_ERR_CAPACITY, "r3d_engine_probe_places: xyz_capacity too small");
    const size_t K1 = d.K + 1;
    DevBuf<double> dbox, dxyz;
    TRY(dbox.alloc(K1 * 5));
    std::vector<ObjBox> hob(1);
    R3D_CUDA(cudaMemcpy(hob.data(), eng->obj.p + object_id, sizeof(ObjBox), cudaMemcpyDeviceToHost));
    const size_t nxyz = xyz_out ? (size_t)nf * hob[0].count * 3 : 0;
    if (nxyz) TRY(dxyz.alloc(nxyz));
    k_probe_points<<<148, 256, 0, st>>>(d, scan, object_id, nf, nxyz ? dxyz.p : nullptr, dbox.p); r3d_count_launch();
    R3D_CUDA(cudaMemcpyAsync(flags_out, eng->cand_flags.p + scan * K1, K1, cudaMemcpyDeviceToHost, st));
    R3D_CUDA(cudaMemcpyAsync(box_out, dbox.p, K1 * 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (nxyz) R3D_CUDA(cudaMemcpyAsync(xyz_out, dxyz.p, nxyz * sizeof(double), cudaMemcpyDeviceToHost, st));
    R3D_CUDA(cudaStreamSynchronize(st));
    return r3d_check_launch("r3d_engine_probe_places");
}
