// Kernels of the batched Real3D-Aug engine.  One "round" = one tried cut object for every unfinished scan of the
// batch; the reference's nested Python loops (od/ins:375-614) are a per-scan state machine advanced on the device
// (k_ctrl), so the host only launches a fixed kernel sequence per round and polls one counter.
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
