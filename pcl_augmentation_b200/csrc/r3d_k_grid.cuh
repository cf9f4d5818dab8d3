// Kernels of the batched Real3D-Aug engine, part: once-per-scan spatial indices (CSR grids, column index, distance transform).
// Included by r3d_engine_kernels.cuh (inside namespace r3d, after the shared constants); not a standalone header.
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
    bool ok = false;                                  // the distinct surface labels of all classes (kernel parameter space)
    for (int j = 0; j < e.n_surf_all; ++j) ok |= lab == (unsigned)e.surf_all[j];
    return ok;
}
// apts.w of original point p with label lab: the point index, and for semseg the slot of the label among the labels
// some class may stand on (EngineDev::surf_all)
__device__ __forceinline__ float apt_tag(const EngineDev& e, unsigned lab, int p) {
    unsigned slot = 0u;
    if (e.task == 1) {
        slot = APT_SLOT_NONE;
        for (int j = 0; j < e.n_surf_all; ++j) if (lab == (unsigned)e.surf_all[j]) slot = (unsigned)j;
    }
    return __uint_as_float((slot << APT_IDX_BITS) | (unsigned)p);
}
__device__ __forceinline__ int grid_coord(const EngineDev& e, float v) {
    const int i = (int)floorf(v * e.grid_inv_cell) + (e.G >> 1);
    return max(0, min(i, e.G - 1));
}

// Chebyshev distance (in cells, capped at NEAR_CAP) from every cell to the nearest cell that holds a surface point
// (k_grid_near_bits, r3d_k_prepass.cuh): the road-level search of a candidate whose surroundings are empty starts at
// that ring instead of growing through the empty ones.
constexpr int NEAR_CAP = 12;

// exclusive prefix sum of the per-cell counts (one CTA per scan, four cells per thread and step); after the scatter
// pass cell[c] = END of cell c
__device__ __forceinline__ void bucket_scan_body(int* arr, size_t stride, int n, int n_scans) {
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
