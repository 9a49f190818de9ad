// Shared host-side plumbing for the C ABI: error text, CUDA checks, launch counting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include <nvtx3/nvToolsExt.h>

#include "../../include/othello_b200.h"

namespace oz {

// Thread-local message behind oz_last_error().
void set_error(const std::string &msg);
int fail(int status, const std::string &msg);

// Number of kernels this library launched (oz_counters::kernel_launches).
extern unsigned long long g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches += (unsigned long long)n; }

// cudaFuncSetAttribute applies to the CURRENT device only, so "already set" is tracked per device ordinal (a process
// may own engines on several GPUs, e.g. MCTS(torch_device="cuda:1") after one on cuda:0).  True the first time it is
// called for `mask` on the current device.
inline bool first_use_on_device(unsigned long long &mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

// NVTX phase ranges (tracing).  Header-only NVTX v3: free unless a profiler is attached, in which case
// the host-side ranges bracket the launches of a phase ("search", "round: select+encode", "round: network", ...).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// Kernel launch with (pdl = true) or without the programmatic-dependent-launch attribute.  A kernel launched with it may
// become resident as soon as every block of its predecessor in the stream has executed griddepcontrol.launch_dependents (or
// exited); it must execute griddepcontrol.wait before it touches anything the predecessor -- or, transitively, anything
// earlier -- wrote.  Both instructions are no-ops in a kernel launched without the attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
