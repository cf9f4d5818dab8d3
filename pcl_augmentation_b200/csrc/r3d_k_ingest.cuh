// Kernels of the batched Real3D-Aug engine, part: A1 + A2 once per original point.
// Included by r3d_engine_kernels.cuh (inside namespace r3d, after the shared constants); not a standalone header.
// ------------------------------------------------------------------------------------------------ ingest
// (the once-per-point ingest itself is k_ingest_count, r3d_k_prepass.cuh)
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
