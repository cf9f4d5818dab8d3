#include "r3d_host.h"
#include "../../include/real3d_b200.h"
#include <atomic>
#include <cstdio>

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

int r3d_fail(int code, const char* msg) { g_last_error = msg; return code; }
int r3d_fail_cuda(cudaError_t err, const char* where) {
    g_last_error = std::string(where) + ": " + cudaGetErrorString(err);
    return R3D_ERR_CUDA;
}
int r3d_check_launch(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return r3d_fail_cuda(e, where);
    return R3D_OK;
}
void r3d_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int r3d_version(void) { return 100; }
extern "C" const char* r3d_last_error(void) { return g_last_error.c_str(); }
extern "C" int64_t r3d_launch_count(void) { return g_launches.load(); }
