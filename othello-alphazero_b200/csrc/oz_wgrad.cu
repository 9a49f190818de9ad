// Weight gradient of the 3x3 convolutions of the training step (reference: the backward of train.py:486-502, i.e.
// torch.autograd of nn.Conv2d in neural_net.py:15,38-45) on tcgen05, without an im2col buffer and without a library GEMM:
//
//     dW[tap][co][ci] = sum over positions p of dZ[p][co] * X[p + tap][ci]            (tap = (dy, dx), zero outside the board)
//
// GEMM view per tap: M = co (128 per CTA), N = ci (128 per CTA, 32 for the first layer), K = positions.  Both operands are
// read straight from the NHWC tensors the forward / input-gradient kernels use, staged by TMA exactly like the forward
// kernel stages its activations: a box of (32 channels, 16 boards, x, 1 board row) lands in shared memory as rows of 64
// bytes, row = (x * 16 + board), SWIZZLE_64B.  Here the 16 boards of one (y, x) square are the K = 16 of one MMA and the 32
// channels of a row are the M / N direction, i.e. the operands are MN-major: the UMMA descriptor of a square is the square's
// 1 KB slot, SBO = 512 B (the next 8 boards), LBO = the distance to the next 32-channel box.  A tap is again nothing but a
// slot offset: X is staged with x = -1 .. 8 (two zero columns from TMA's out-of-bounds fill), so tap dx of square x is slot
// x + dx + 1, and tap dy is the board row the box was taken from.
//
// One CTA owns one dy (three taps = three 128-column fp32 accumulators in TMEM, 384 of 512 columns), one 128 x 128 tile of
// (co, ci) and a slice of the boards (split K); it streams (unit of 16 boards, board row) passes through a 3-stage ring
// (72 KB per stage: 4 + 4 boxes), 24 MMAs of 128 x 128 x 16 per pass, and adds its accumulators into G with vector
// reductions at the end.  Roles: warp 0 TMA, warp 1 MMA issue, warps 2-5 epilogue.
//
// Roofline: tensor-bound in FLOPs (2 * 9 * co * ci per position) but every K = 16 slab of dZ and X is used by three MMAs
// only, so the kernel moves 136 KB from L2 per 25 MFLOP (185 FLOP/B) and single-CTA 128 x 128 x 16 MMAs read 8 KB of shared
// memory per 64 clocks (the 128 B/clk ceiling): expect about half of the dense tensor peak, not the forward kernel's 0.95.
#include <cuda_bf16.h>

#include <algorithm>
#include <string>

#include "oz_resnet.cuh"
#include "oz_tc.cuh"

namespace oz {
namespace wg {
constexpr int kStages = 3;
constexpr int kSlot = 1024;                       // one (y, x) square: 16 boards x 32 channels x 2 B
constexpr int kABox = 8 * kSlot;                  // dZ: x = 0 .. 7
constexpr int kBBox = 10 * kSlot;                 // X:  x = -1 .. 8
constexpr int kStageBytes = 4 * kABox + 4 * kBBox;
constexpr int kMisc = 1024;
constexpr int kSmem = 1024 + kStages * kStageBytes + kMisc;
constexpr int kThreads = 6 * 32;
static_assert(kSmem <= 232448, "shared memory budget");
}  // namespace wg

struct WgradParams {
    float *G;            // [cout][9][cinp] fp32, zeroed by the caller; partial sums are added
    int units;           // boards / 16
    int mtiles, ntiles;  // cout / 128, max(1, cinp / 128)
    int nblocks;         // 32-channel blocks of X per tile: 4, or 1 for the first layer (cinp = 32)
    int splits;          // CTAs sharing the boards of one (dy, tile)
    int cinp;
};

// MN-major, SWIZZLE_64B shared-memory matrix descriptor: rows of 64 B = 32 elements of the M / N direction, 8 rows (K) per
// 512-byte atom; SBO = next 8 K rows, LBO = next 32 M / N elements.
__device__ __forceinline__ uint64_t umma_desc_mn_sw64(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_mma_1cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_1cta(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *ptr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(wg::kThreads, 1)
wgrad3x3_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x, const WgradParams p) {
    using namespace wg;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t misc = base + kStages * kStageBytes;
    const uint32_t bar_full = misc, bar_empty = misc + 8 * kStages, bar_done = misc + 16 * kStages;
    uint32_t *tmem_ptr = (uint32_t *)(gen + kStages * kStageBytes + 128);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // work item of this CTA: (dy, co tile, ci tile, split)
    int w = blockIdx.x;
    const int split = w % p.splits; w /= p.splits;
    const int nt = w % p.ntiles; w /= p.ntiles;
    const int mt = w % p.mtiles; w /= p.mtiles;
    const int dyi = w, dy = dyi - 1;
    const int u0 = (int)((long)p.units * split / p.splits), u1 = (int)((long)p.units * (split + 1) / p.splits);
    const int y_lo = dy < 0 ? 1 : 0, y_hi = dy > 0 ? 7 : 8;           // board rows whose tap row y + dy is on the board
    const int ny = y_hi - y_lo, np = (u1 - u0) * ny;
    const int ncols = p.nblocks * 32;                                  // N of one MMA = columns of one accumulator
    const uint32_t stage_tx = (uint32_t)(4 * kABox + p.nblocks * kBBox);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dz) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    griddep_launch_dependents();

    if (warp == 0) {
        // ===== TMA producer: per pass four 32-channel boxes of dZ (row y) and of X (row y + dy, x = -1 .. 8) =====
        griddep_wait();                                   // dZ and X come from the kernels in front of this one
        for (int it = 0; it < np; ++it) {
            const int s = it % kStages, ph = (it / kStages) & 1;
            const int unit = u0 + it / ny, y = y_lo + it % ny;
            mbar_wait_b(bar_empty + 8 * s, ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8 * s, stage_tx);
                const uint32_t a0 = base + s * kStageBytes, b0 = a0 + 4 * kABox;
                for (int j = 0; j < 4; ++j) tma_load_4d(a0 + j * kABox, &map_dz, bar_full + 8 * s, mt * 128 + j * 32, unit * 16, 0, y);
                for (int j = 0; j < p.nblocks; ++j) tma_load_4d(b0 + j * kBBox, &map_x, bar_full + 8 * s, nt * 128 + j * 32, unit * 16, -1, y + dy);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[dx] (co x ci) += dZ(square x)^T . X(square x + dx), both operands MN-major =====
        // kind::f16: bf16 x bf16 -> f32, A and B MN-major (bits 15, 16), M = 128, N = ncols
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(ncols >> 3) << 17) | ((128u >> 4) << 24);
        for (int it = 0; it < np; ++it) {
            const int s = it % kStages, ph = (it / kStages) & 1;
            mbar_wait_b(bar_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t a0 = base + s * kStageBytes, b0 = a0 + 4 * kABox;
            if (elect_one()) {
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const uint64_t adesc = umma_desc_mn_sw64(a0 + x * kSlot, kABox);
#pragma unroll
                    for (int dxi = 0; dxi < 3; ++dxi)
                        tc_mma_1cta(tmem_base + dxi * ncols, adesc, umma_desc_mn_sw64(b0 + (x + dxi) * kSlot, kBBox), idesc, (it | x) != 0);
                }
                tc_commit_1cta(bar_empty + 8 * s);
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit_1cta(bar_done);
        __syncwarp();
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter (warp % 4) = 32 co rows; add the accumulators into G =====
        const int quad = warp & 3, co = mt * 128 + quad * 32 + lane;
        if (np > 0) {
            mbar_wait_b(bar_done, 0);
            tc_fence_after();
            for (int dxi = 0; dxi < 3; ++dxi) {
                float *row = p.G + ((size_t)co * 9 + (dyi * 3 + dxi)) * p.cinp + nt * 128;
                for (int c = 0; c < p.nblocks; ++c) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + dxi * ncols + c * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        red_add_v4(row + c * 32 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int launch_wgrad3x3(const void *d_x, int cinp, const void *d_dz, int cout, size_t boards, float *d_G, cudaStream_t stream) {
    if (!(cinp == 32 || cinp % 128 == 0) || cout % 128 != 0 || boards % 16 != 0 || boards == 0)
        return fail(OZ_ERR_INVALID_ARGUMENT, "wgrad3x3: cinp must be 32 or a multiple of 128, cout a multiple of 128, boards a multiple of 16");
    static unsigned long long attr_devices = 0;
    if (first_use_on_device(attr_devices))
        OZ_CUDA(cudaFuncSetAttribute(wgrad3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::kSmem));
    CUtensorMap mdz, mx;
    OZ_TRY(make_act_map(&mdz, const_cast<void *>(d_dz), boards, cout, 1, 8));
    OZ_TRY(make_act_map(&mx, const_cast<void *>(d_x), boards, cinp, 1, 10));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    WgradParams p{};
    p.G = d_G; p.units = (int)(boards / 16); p.mtiles = cout / 128; p.ntiles = std::max(1, cinp / 128);
    p.nblocks = cinp == 32 ? 1 : 4; p.cinp = cinp;
    const int tiles = 3 * p.mtiles * p.ntiles;
    p.splits = std::max(1, std::min(p.units, sms / tiles));
    OZ_CUDA(cudaMemsetAsync(d_G, 0, (size_t)cout * 9 * cinp * sizeof(float), stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(tiles * p.splits)); cfg.blockDim = dim3(wg::kThreads); cfg.dynamicSmemBytes = wg::kSmem; cfg.stream = stream;
    OZ_CUDA(cudaLaunchKernelEx(&cfg, wgrad3x3_kernel, mdz, mx, p));
    count_launch();
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

}  // namespace oz
