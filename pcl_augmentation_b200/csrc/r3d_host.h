// Host-side helpers shared by the translation units of libreal3d_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <string>

int r3d_fail(int code, const char* msg);
int r3d_fail_cuda(cudaError_t err, const char* where);
int r3d_check_launch(const char* where);   // cudaGetLastError() -> R3D_OK / R3D_ERR_CUDA
void r3d_count_launch(int n = 1);

#define R3D_CUDA(call)                                                    \
    do {                                                                  \
        cudaError_t _e = (call);                                          \
        if (_e != cudaSuccess) return r3d_fail_cuda(_e, #call);           \
    } while (0)
