// Kernels of the batched Real3D-Aug engine, one header per stage.  Two execution models share them:
//  * the per-scan walker (default; r3d_k_walk.cuh): streaming kernels once per scan (r3d_k_prepass.cuh, r3d_closefill.cuh,
//    r3d_k_output.cuh) around ONE persistent CTA per scan that runs the reference's whole loop (od/ins:375-614);
//  * the staged rounds of round 1 (r3d_k_ctrl / _update_project / _placement / _occlusion as kernels): one "round" = one
//    tried cut object for every unfinished scan, the per-scan state machine advanced on the device (k_ctrl), the host
//    launching a fixed kernel sequence per round.  Kept because it evaluates EVERY candidate of a try (debug_candidates).
#pragma once
#include "r3d_engine.cuh"
#include "r3d_closefill.cuh"

namespace r3d {

constexpr int CHUNK = 4096;          // points per CTA of the streaming kernels (256 threads x 16 points)
constexpr int STREAM_THREADS = 256;

__device__ __forceinline__ void set_error(ScanState& s, int code) {
    if (s.status == 0) s.status = code;
}


#include "r3d_k_ingest.cuh"
#include "r3d_k_ctrl.cuh"
#include "r3d_k_update_project.cuh"
#include "r3d_k_grid.cuh"
#include "r3d_k_prepass.cuh"
#include "r3d_k_placement.cuh"
#include "r3d_k_occlusion.cuh"
#include "r3d_k_output.cuh"
// the per-scan walker in its two CTA shapes (see the head of r3d_k_walk.cuh)
namespace walk_od {
#ifndef R3D_WALK_OD_THREADS
#define R3D_WALK_OD_THREADS 384
#endif
#ifndef R3D_WALK_OD_CTAS
#define R3D_WALK_OD_CTAS 3
#endif
#define R3D_WALK_THREADS R3D_WALK_OD_THREADS
#define R3D_WALK_CTAS_PER_SM R3D_WALK_OD_CTAS
#include "r3d_k_walk.cuh"
#undef R3D_WALK_THREADS
#undef R3D_WALK_CTAS_PER_SM
}  // namespace walk_od
namespace walk_ss {
#define R3D_WALK_THREADS 512
#define R3D_WALK_CTAS_PER_SM 2
#include "r3d_k_walk.cuh"
#undef R3D_WALK_THREADS
#undef R3D_WALK_CTAS_PER_SM
}  // namespace walk_ss

}  // namespace r3d
