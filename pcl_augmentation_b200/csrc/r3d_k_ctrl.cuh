// Kernels of the batched Real3D-Aug engine, part: A13: the per-scan scheduling state machine.
// Included by r3d_engine_kernels.cuh (inside namespace r3d, after the shared constants); not a standalone header.
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

// one step of the state machine for scan b (ONE thread): what the scan needs next — apply the last vis_px mask,
// refresh the range image (new slot after a scene change), try a cut object.  Shared by the staged round kernel k_ctrl
// and the per-scan persistent walker (r3d_k_walk.cuh).
__device__ void ctrl_advance(const EngineDev& e, int b, ScanState& s, int& apply, int& project, int& tryact) {
    apply = 0; project = 0; tryact = 0;
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
}

__global__ void __launch_bounds__(32) k_ctrl(EngineDev e, int n_scans) {
    const int b = blockIdx.x;
    if (b >= n_scans) return;
    ScanState& s = e.st[b];
    if (threadIdx.x == 0) {
        const unsigned round = *(volatile unsigned*)&e.round_ctl[0];
        const int slot = (int)(round & 63u);
        int apply = 0, project = 0, tryact = 0;
        ctrl_advance(e, b, s, apply, project, tryact);
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
