// Shared host-side plumbing for the C ABI: error text, CUDA checks, launch counting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/othello_b200.h"

namespace oz {

// Thread-local message behind oz_last_error().
void set_error(const std::string &msg);
int fail(int status, const std::string &msg);

// Number of kernels this library launched (oz_counters::kernel_launches).
extern unsigned long long g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches += (unsigned long long)n; }

}  // namespace oz

#define OZ_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return oz::fail(OZ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define OZ_TRY(expr)                     \
    do {                                 \
        int _s = (expr);                 \
        if (_s != OZ_OK) return _s;      \
    } while (0)
