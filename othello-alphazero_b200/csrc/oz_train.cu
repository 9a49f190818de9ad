// Native training step of the AlphaZero ResNet.
// reference: python/othello_alphazero/train.py:486-502 (the body of the batch loop of _train: forward in train mode,
// policy cross-entropy + value MSE + explicit L2 over all parameters, backward, SGD with momentum) on the network of
// python/othello_alphazero/neural_net.py:9-172.
//
// The 3x3 convolutions -- forward and input gradients -- run on the tcgen05 implicit-GEMM kernel of oz_resnet.cu in
// bf16 with fp32 accumulation (the input gradient of a convolution is the same convolution with the taps rotated by 180
// degrees and the channel roles swapped, so only the weight packing differs).  Weight gradients run on the tcgen05 kernel
// of oz_wgrad.cu (MN-major operands staged from the same NHWC tensors, nine taps as slot offsets, split K over the boards):
// no im2col buffer and no library GEMM -- every kernel of the step is this project's.  BatchNorm uses batch statistics (train mode) in fp32; the heads, the losses and the optimizer are fp32.  Master
// weights, gradients and momentum are fp32 vectors in the order of AlphaZeroNet.parameters().
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "oz_common.cuh"
#include "oz_resnet.cuh"

namespace oz {
namespace {

typedef __nv_bfloat16 bf16;

// Programmatic dependent launch for the many small kernels of a step: every kernel lets its successor become resident at
// once and waits for its predecessor's results before touching memory, so the launch latency of a kernel overlaps the
// execution of the one before it (a kernel completes only after its own wait, hence completion stays transitive).
#define OZ_PDL_PROLOGUE() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    count_launch();
}

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of `v` (blockDim.x <= 1024); result valid in thread 0.
__device__ float block_sum(float v, float *scratch /* >= 32 floats */) {
    v = warp_sum_f(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x + 31) / 32 ? scratch[threadIdx.x] : 0.f;
        r = warp_sum_f(r);
    }
    return r;
}

// ------------------------------------------------------------------ weight packing
// master fp32 W [C][cin][9] (PyTorch OIHW, tap = ky*3+kx) ->
//   wf bf16 [9][C][cinp]          forward operand of conv3x3_pair_kernel ([tap][cout][cin], cin zero-padded to cinp)
//   wd bf16 [9][cin][C] (cin==C)  input-gradient operand: Wd[8 - tap][ci][co] = W[co][ci][tap]
struct PackDesc { size_t w; int cin, cinp; bf16 *wf, *wd; };      // one per conv layer (blockIdx.y)
__global__ void pack_weights_kernel(const float *__restrict__ params, const PackDesc *__restrict__ desc, int C) {
    OZ_PDL_PROLOGUE();
    const PackDesc d = desc[blockIdx.y];
    const float *w = params + d.w;
    const int cin = d.cin, cinp = d.cinp;
    bf16 *wf = d.wf, *wd = d.wd;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 9 * C * cinp) return;
    const int ci = idx % cinp, co = (idx / cinp) % C, tap = idx / (cinp * C);
    const float v = ci < cin ? w[((size_t)co * cin + ci) * 9 + tap] : 0.f;
    wf[idx] = __float2bfloat16(v);
    if (wd && ci < cin) wd[((size_t)(8 - tap) * cin + ci) * C + co] = __float2bfloat16(v);
}

// ------------------------------------------------------------------ BatchNorm (train mode), tensors bf16 [P][C]
// All four kernels move 16 bytes (8 channels) per thread and access.
__device__ __forceinline__ void unpack8(const uint4 &q, float (&f)[8]) {
    const __nv_bfloat162 *h = (const __nv_bfloat162 *)&q;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 v = __bfloat1622float2(h[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 q;
    __nv_bfloat162 *h = (__nv_bfloat162 *)&q;
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return q;
}

// Shared tail of the two reductions: 16 per-thread partial sums (8 channels x 2 statistics) -> one atomic per channel,
// statistic and block.  out[copy][c] += first statistic, out[copy][C + c] += second, copy = block % kBnCopies: a few hundred
// blocks adding to the same 2C addresses serialise in L2 (about 50 ns each), which at one copy was the whole run time of the
// reductions; consumers add the copies up (bn_total).
constexpr int kBnCopies = 16;
__device__ __forceinline__ void bn_block_reduce(float (&a)[8], float (&b)[8], int C, int tpr, int rows, int cg, int r0, float *out) {
    __shared__ float red[16][256];
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[i][threadIdx.x] = a[i]; red[8 + i][threadIdx.x] = b[i]; }
    __syncthreads();
    if (r0 == 0) {
        for (int r = 1; r < rows; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] += red[i][r * tpr + cg]; b[i] += red[8 + i][r * tpr + cg]; }
        out += (size_t)(blockIdx.x % kBnCopies) * 2 * C;     // one of kBnCopies accumulators: same-address atomics serialise in L2
#pragma unroll
        for (int i = 0; i < 8; ++i) { atomicAdd(&out[8 * cg + i], a[i]); atomicAdd(&out[C + 8 * cg + i], b[i]); }
    }
}
// Total of the kBnCopies partial accumulators of statistic slot `idx` (0 .. 2C-1).
__device__ __forceinline__ float bn_total(const float *sums, int C, int idx) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kBnCopies; ++k) t += sums[(size_t)k * 2 * C + idx];
    return t;
}

// The four BatchNorm passes are pure HBM streams (16 bytes per thread and load).  A thread owns one group of 8 channels
// (cg) and walks board squares p = first, first + stride, ...: its per-channel constants (scale / shift, mean / invstd, the
// reduction terms) live in registers for the whole kernel, and kBnUnroll independent 16-byte loads per tensor are in flight
// before the first is used -- with one load in flight per thread and the constants re-derived per element (two divisions
// and an rsqrt per channel) the passes ran at 40 % of the HBM bandwidth and were instruction-bound.
constexpr int kBnUnroll = 4;

// Block-cooperative version for the streaming consumers: thread i adds up the copies of statistic slots i, i + blockDim, ...
// into shared memory once per block (tot[2C]); 2 * kBnCopies loads per thread instead of 16 * kBnCopies.
__device__ __forceinline__ void bn_totals_to_smem(const float *sums, int C, float *tot) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) tot[i] = bn_total(sums, C, i);
    __syncthreads();
}

// mean / invstd of channel c from the sums (every consumer recomputes them: two flops instead of a kernel launch).
__device__ __forceinline__ void bn_moments_tot(const float *tot, int C, int c, float count, float eps, float &mean, float &invstd, float &var) {
    mean = tot[c] / count;
    var = fmaxf(tot[C + c] / count - mean * mean, 0.f);
    invstd = rsqrtf(var + eps);
}

// a = relu(gamma * (z - mean) * invstd + beta [+ skip]).  Block 0 also stores mean / invstd for the backward pass and
// updates the running statistics as nn.BatchNorm2d does (momentum, unbiased variance).  The convolution bias only shifts
// the mean: the normalised output does not depend on it, the running mean does.
__global__ void __launch_bounds__(256) bn_apply_kernel(const bf16 *__restrict__ z, const bf16 *__restrict__ skip, size_t P, int C,
                                                       const float *__restrict__ sums, const float *__restrict__ gamma, const float *__restrict__ beta,
                                                       const float *__restrict__ conv_bias, float eps, float momentum, float *__restrict__ mean_out,
                                                       float *__restrict__ invstd_out, float *__restrict__ run_mean, float *__restrict__ run_var,
                                                       bf16 *__restrict__ a) {
    OZ_PDL_PROLOGUE();
    __shared__ float tot[2 * 384];
    bn_totals_to_smem(sums, C, tot);
    const float count = (float)P;
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float m, inv, var;
            bn_moments_tot(tot, C, c, count, eps, m, inv, var);
            mean_out[c] = m; invstd_out[c] = inv;
            run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (m + conv_bias[c]);
            run_var[c] = (1.f - momentum) * run_var[c] + momentum * var * (count / fmaxf(count - 1.f, 1.f));
        }
    const int tpr = C / 8, rows = blockDim.x / tpr;
    const int cg = threadIdx.x % tpr, r0 = threadIdx.x / tpr;
    if (r0 >= rows) return;
    float g[8], m[8], iv[8], bt[8];                       // y = g * (z - m) * iv + bt, evaluated in this order (as before)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float var;
        bn_moments_tot(tot, C, 8 * cg + k, count, eps, m[k], iv[k], var);
        g[k] = gamma[8 * cg + k]; bt[k] = beta[8 * cg + k];
    }
    const size_t stride = (size_t)gridDim.x * rows;
    for (size_t p = (size_t)blockIdx.x * rows + r0; p < P; p += kBnUnroll * stride) {
        uint4 zq[kBnUnroll], sq[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const size_t pu = p + u * stride;
            if (pu < P) { zq[u] = *(const uint4 *)(z + pu * C + 8 * cg); if (skip) sq[u] = *(const uint4 *)(skip + pu * C + 8 * cg); }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const size_t pu = p + u * stride;
            if (pu >= P) break;
            float v[8], sk[8];
            unpack8(zq[u], v);
            if (skip) unpack8(sq[u], sk);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float y = g[k] * (v[k] - m[k]) * iv[k] + bt[k];
                if (skip) y += sk[k];
                v[k] = fmaxf(y, 0.f);
            }
            *(uint4 *)(a + pu * C + 8 * cg) = pack8(v);
        }
    }
}

// Backward through ReLU and BatchNorm.  dy = da * [a > 0]; s[c] += sum dy, s[C + c] += sum dy * zhat.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const bf16 *__restrict__ da, const bf16 *__restrict__ a, const bf16 *__restrict__ z,
                                                            size_t P, int C, const float *__restrict__ mean, const float *__restrict__ invstd,
                                                            float *__restrict__ s) {
    OZ_PDL_PROLOGUE();
    const int tpr = C / 8, rows = blockDim.x / tpr;
    const int cg = threadIdx.x % tpr, r0 = threadIdx.x / tpr;
    float m[8], iv[8], s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { m[i] = mean[8 * cg + i]; iv[i] = invstd[8 * cg + i]; s1[i] = 0.f; s2[i] = 0.f; }
    if (r0 < rows) {
        const size_t stride = (size_t)gridDim.x * rows;
        for (size_t p = (size_t)blockIdx.x * rows + r0; p < P; p += kBnUnroll * stride) {
            uint4 gq[kBnUnroll], aq[kBnUnroll], zq[kBnUnroll];
#pragma unroll
            for (int u = 0; u < kBnUnroll; ++u) {
                const size_t pu = p + u * stride, o = pu * C + 8 * cg;
                if (pu < P) { gq[u] = *(const uint4 *)(da + o); aq[u] = *(const uint4 *)(a + o); zq[u] = *(const uint4 *)(z + o); }
            }
#pragma unroll
            for (int u = 0; u < kBnUnroll; ++u) {
                if (p + u * stride >= P) break;
                float g[8], av[8], zv[8];
                unpack8(gq[u], g); unpack8(aq[u], av); unpack8(zq[u], zv);
#pragma unroll
                for (int i = 0; i < 8; ++i) { const float d = av[i] > 0.f ? g[i] : 0.f; s1[i] += d; s2[i] += d * (zv[i] - m[i]) * iv[i]; }
            }
        }
    }
    bn_block_reduce(s1, s2, C, tpr, rows, cg, r0 < rows ? r0 : 1 << 30, s);
}

// dz = gamma * invstd * (dy - s1 / M - zhat * s2 / M); optionally keeps dy (the gradient that also flows into the skip
// connection).  Block 0 writes dgamma = s2, dbeta = s1.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const bf16 *__restrict__ da, const bf16 *__restrict__ a, const bf16 *__restrict__ z,
                                                           size_t P, int C, const float *__restrict__ mean, const float *__restrict__ invstd,
                                                           const float *__restrict__ gamma, const float *__restrict__ s, bf16 *__restrict__ dz,
                                                           bf16 *__restrict__ dy_out, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    OZ_PDL_PROLOGUE();
    __shared__ float tot[2 * 384];
    bn_totals_to_smem(s, C, tot);
    if (blockIdx.x == 0) for (int c = threadIdx.x; c < C; c += blockDim.x) { dgamma[c] = tot[C + c]; dbeta[c] = tot[c]; }
    const int tpr = C / 8, rows = blockDim.x / tpr;
    const int cg = threadIdx.x % tpr, r0 = threadIdx.x / tpr;
    if (r0 >= rows) return;
    const float invM = 1.f / (float)P;
    float m[8], iv[8], gi[8], t1[8], t2[8];               // dz = gi * (d - t1 - zh * t2)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = 8 * cg + k;
        m[k] = mean[c]; iv[k] = invstd[c]; gi[k] = gamma[c] * invstd[c]; t1[k] = tot[c] * invM; t2[k] = tot[C + c] * invM;
    }
    const size_t stride = (size_t)gridDim.x * rows;
    for (size_t p = (size_t)blockIdx.x * rows + r0; p < P; p += kBnUnroll * stride) {
        uint4 gq[kBnUnroll], aq[kBnUnroll], zq[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const size_t pu = p + u * stride, o = pu * C + 8 * cg;
            if (pu < P) { gq[u] = *(const uint4 *)(da + o); aq[u] = *(const uint4 *)(a + o); zq[u] = *(const uint4 *)(z + o); }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const size_t pu = p + u * stride, o = pu * C + 8 * cg;
            if (pu >= P) break;
            float g[8], av[8], zv[8], d[8];
            unpack8(gq[u], g); unpack8(aq[u], av); unpack8(zq[u], zv);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                d[k] = av[k] > 0.f ? g[k] : 0.f;
                const float zh = (zv[k] - m[k]) * iv[k];
                g[k] = gi[k] * (d[k] - t1[k] - zh * t2[k]);
            }
            *(uint4 *)(dz + o) = pack8(g);
            if (dy_out) *(uint4 *)(dy_out + o) = pack8(d);
        }
    }
}

// ------------------------------------------------------------------ weight gradient: wgrad3x3_kernel (oz_wgrad.cu) + scatter
// G fp32 [C][9][cinp] (output of wgrad3x3_kernel) -> dW fp32 [C][cin][9] (OIHW)
__global__ void wgrad_scatter_kernel(const float *__restrict__ G, int C, int cin, int cinp, float *__restrict__ dw) {
    OZ_PDL_PROLOGUE();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * cin * 9) return;
    const int tap = idx % 9, ci = (idx / 9) % cin, co = idx / (9 * cin);
    dw[idx] = G[((size_t)co * 9 + tap) * cinp + ci];
}

// ------------------------------------------------------------------ heads (fp32)
// h[k][p] = sum_c a[p][c] * w[k][c] + b[k], k = 0, 1 (policy), 2 (value); one warp per position.
__global__ void head_conv_fwd_kernel(const bf16 *__restrict__ a, size_t P, int C, const float *__restrict__ pw, const float *__restrict__ pb,
                                     const float *__restrict__ vw, const float *__restrict__ vb, float *__restrict__ h) {
    OZ_PDL_PROLOGUE();
    const size_t p = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float v = __bfloat162float(a[p * C + c]);
        d0 += v * pw[c]; d1 += v * pw[C + c]; d2 += v * vw[c];
    }
    d0 = warp_sum_f(d0); d1 = warp_sum_f(d1); d2 = warp_sum_f(d2);
    if (lane == 0) { h[p] = d0 + pb[0]; h[P + p] = d1 + pb[1]; h[2 * P + p] = d2 + vb[0]; }
}

// BatchNorm + ReLU of the three head channels (blockIdx.y = channel k): hn[k][p] = relu(bn(h[k][p])); also scattered into
// the flattened FC inputs hp [N][128] (index k*64 + pos) and hv [N][64].  stat[k] = mean, stat[3 + k] = invstd.  Two kernels,
// many blocks each (one block per channel took 0.2 ms at batch 4 096): acc[2k] += sum h, acc[2k + 1] += sum h^2 (zeroed
// per step), then every block derives the moments from acc.
__global__ void __launch_bounds__(256) head_bn_stats_kernel(const float *__restrict__ h, size_t P, float *__restrict__ acc) {
    OZ_PDL_PROLOGUE();
    __shared__ float scratch[32];
    const int k = blockIdx.y;
    const float *hk = h + (size_t)k * P;
    float s = 0.f, q = 0.f;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) { const float v = hk[p]; s += v; q += v * v; }
    s = block_sum(s, scratch); q = block_sum(q, scratch);
    if (threadIdx.x == 0) { atomicAdd(&acc[2 * k], s); atomicAdd(&acc[2 * k + 1], q); }
}

__global__ void __launch_bounds__(256) head_bn_fwd_kernel(const float *__restrict__ h, size_t P, const float *__restrict__ acc,
                                   const float *__restrict__ gamma_p, const float *__restrict__ beta_p,
                                   const float *__restrict__ gamma_v, const float *__restrict__ beta_v, float eps, float momentum,
                                   float *__restrict__ run_p /* mean[2], var[2] */, float *__restrict__ run_v /* mean[1], var[1] */,
                                   float *__restrict__ stat, float *__restrict__ hn, float *__restrict__ hp, float *__restrict__ hv) {
    OZ_PDL_PROLOGUE();
    const int k = blockIdx.y;
    const float *hk = h + (size_t)k * P;
    const float cnt = (float)P, mean = acc[2 * k] / cnt, var = fmaxf(acc[2 * k + 1] / cnt - mean * mean, 0.f), inv = rsqrtf(var + eps);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        stat[k] = mean; stat[3 + k] = inv;
        float *rm = k < 2 ? run_p + k : run_v, *rv = k < 2 ? run_p + 2 + k : run_v + 1;
        *rm = (1.f - momentum) * *rm + momentum * mean;
        *rv = (1.f - momentum) * *rv + momentum * var * (cnt / fmaxf(cnt - 1.f, 1.f));
    }
    const float g = k < 2 ? gamma_p[k] : gamma_v[0], b = k < 2 ? beta_p[k] : beta_v[0];
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
        const float y = fmaxf(g * (hk[p] - mean) * inv + b, 0.f);
        hn[(size_t)k * P + p] = y;
        const size_t n = p >> 6; const int pos = (int)(p & 63);
        if (k < 2) hp[n * 128 + k * 64 + pos] = y; else hv[n * 64 + pos] = y;
    }
}

// One warp per board: policy logits -> softmax -> loss and dlogits; value MLP -> tanh -> loss and its gradients.
// loss[0] += -sum_j t_j log p_j / N, loss[1] += (v - t)^2 / N.
constexpr int kFcLwStride = 132, kFcW1Stride = 68;      // floats; = 4 mod 32: conflict-free 128-bit loads when lanes read different rows
__global__ void __launch_bounds__(256) head_fc_fwd_kernel(const float *__restrict__ hp, const float *__restrict__ hv, int N, int Hv, const float *__restrict__ lw,
                                   const float *__restrict__ lb, const float *__restrict__ w1, const float *__restrict__ b1,
                                   const float *__restrict__ w2, const float *__restrict__ b2, const float *__restrict__ tp,
                                   const float *__restrict__ tv, float *__restrict__ policy, float *__restrict__ value, float *__restrict__ dlogits,
                                   float *__restrict__ u, float *__restrict__ dvpre, float *__restrict__ loss) {
    OZ_PDL_PROLOGUE();
    // FC weights staged in shared memory once per block; a warp takes one board at a time and a LANE owns outputs (policy
    // logits lane and lane + 32, hidden units lane + 32 j): every weight row is walked with 128-bit loads against the board's
    // inputs broadcast from shared memory -- no warp reduction per output (the first version did one per logit and hidden unit:
    // 193 five-step butterflies per board, 0.12 ms at batch 4 096).  Summation order per output: i = 0, 1, 2, ...
    extern __shared__ __align__(16) float fc_sm[];
    float *s_lw = fc_sm, *s_w1 = s_lw + 65 * kFcLwStride, *s_x = s_w1 + Hv * kFcW1Stride;      // s_x: [warps][192] inputs of the warp's board
    for (int i = threadIdx.x; i < 65 * 128; i += blockDim.x) s_lw[(i >> 7) * kFcLwStride + (i & 127)] = lw[i];
    for (int i = threadIdx.x; i < Hv * 64; i += blockDim.x) s_w1[(i >> 6) * kFcW1Stride + (i & 63)] = w1[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warps = blockDim.x >> 5, wib = threadIdx.x >> 5;
    float *x = s_x + wib * 192;
    float loss_p = 0.f, loss_v = 0.f;              // lane 0: this warp's share of the two losses (one atomic pair per BLOCK at the end)
    for (int n = blockIdx.x * warps + wib; n < N; n += gridDim.x * warps) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) x[lane + 32 * k] = hp[(size_t)n * 128 + lane + 32 * k];
#pragma unroll
        for (int k = 0; k < 2; ++k) x[128 + lane + 32 * k] = hv[(size_t)n * 64 + lane + 32 * k];
        __syncwarp();
        // ---- policy logits: lane, lane + 32; logit 64 as one warp reduction
        float lg[3];
        {
            const float4 *r0 = (const float4 *)(s_lw + lane * kFcLwStride), *r1 = (const float4 *)(s_lw + (lane + 32) * kFcLwStride);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
            for (int i4 = 0; i4 < 32; ++i4) {
                const float4 f = *(const float4 *)(x + 4 * i4), p0 = r0[i4], p1 = r1[i4];
                a0 += f.x * p0.x; a0 += f.y * p0.y; a0 += f.z * p0.z; a0 += f.w * p0.w;
                a1 += f.x * p1.x; a1 += f.y * p1.y; a1 += f.z * p1.z; a1 += f.w * p1.w;
            }
            const float4 f = *(const float4 *)(x + 4 * lane), p2 = *(const float4 *)(s_lw + 64 * kFcLwStride + 4 * lane);
            float a2 = f.x * p2.x; a2 += f.y * p2.y; a2 += f.z * p2.z; a2 += f.w * p2.w;
            a2 = warp_sum_f(a2);
            lg[0] = a0 + lb[lane]; lg[1] = a1 + lb[lane + 32]; lg[2] = lane == 0 ? a2 + lb[64] : -INFINITY;
        }
        float m = fmaxf(fmaxf(lg[0], lg[1]), lg[2]);
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float e[3], se = 0.f, st = 0.f, ce = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) { const int j = lane + 32 * r; e[r] = j < 65 ? expf(lg[r] - m) : 0.f; se += e[r]; if (j < 65) st += tp[(size_t)n * 65 + j]; }
        se = warp_sum_f(se); st = warp_sum_f(st);
        const float lse = logf(se);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int j = lane + 32 * r;
            if (j < 65) {
                const float p = e[r] / se, t = tp[(size_t)n * 65 + j];
                policy[(size_t)n * 65 + j] = p;
                ce -= t * (lg[r] - m - lse);                        // -t log p
                dlogits[(size_t)n * 65 + j] = (p * st - t) / (float)N;
            }
        }
        ce = warp_sum_f(ce);
        // ---- value head: hidden units lane + 32 j
        float acc2 = 0.f;
        for (int k = lane; k < Hv; k += 32) {
            const float4 *r = (const float4 *)(s_w1 + k * kFcW1Stride);
            float a = 0.f;
#pragma unroll 8
            for (int i4 = 0; i4 < 16; ++i4) {
                const float4 f = *(const float4 *)(x + 128 + 4 * i4), w = r[i4];
                a += f.x * w.x; a += f.y * w.y; a += f.z * w.z; a += f.w * w.w;
            }
            a = fmaxf(a + b1[k], 0.f);
            u[(size_t)n * Hv + k] = a;
            acc2 += a * w2[k];
        }
        acc2 = warp_sum_f(acc2);
        if (lane == 0) {
            const float v = tanhf(acc2 + b2[0]), d = v - tv[n];
            value[n] = v;
            dvpre[n] = 2.f * d / (float)N * (1.f - v * v);
            loss_p += ce / (float)N;
            loss_v += d * d / (float)N;
        }
    }
    __syncthreads();                               // the staged weights are dead: reuse the first floats for the block's sums
    if (lane == 0) { fc_sm[2 * wib] = loss_p; fc_sm[2 * wib + 1] = loss_v; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
        for (int w = 0; w < warps; ++w) t += fc_sm[2 * w + threadIdx.x];
        atomicAdd(&loss[threadIdx.x], t);
    }
}

// Gradients of the FC parameters: one thread per output element, loops over the batch.
__global__ void head_fc_bwd_params_kernel(const float *__restrict__ hp, const float *__restrict__ hv, const float *__restrict__ dlogits,
                                          const float *__restrict__ u, const float *__restrict__ dvpre, const float *__restrict__ w2, int N, int Hv,
                                          float *__restrict__ dlw, float *__restrict__ dlb, float *__restrict__ dw1, float *__restrict__ db1,
                                          float *__restrict__ dw2, float *__restrict__ db2) {
    OZ_PDL_PROLOGUE();
    // blockIdx.y = slice of the batch: every thread sums its parameter's products over one slice and adds the partial sum
    // into the (zeroed) gradient -- one thread walking all N samples serially cost 0.85 ms at batch 4 096.
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n0 = (int)((long)N * blockIdx.y / gridDim.y), n1 = (int)((long)N * (blockIdx.y + 1) / gridDim.y);
    const int n_lw = 65 * 128, n_w1 = Hv * 64;
    if (idx < n_lw) {
        const int j = idx / 128, i = idx % 128;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int n = n0;
        for (; n + 3 < n1; n += 4) {
            a0 += dlogits[(size_t)n * 65 + j] * hp[(size_t)n * 128 + i];
            a1 += dlogits[(size_t)(n + 1) * 65 + j] * hp[(size_t)(n + 1) * 128 + i];
            a2 += dlogits[(size_t)(n + 2) * 65 + j] * hp[(size_t)(n + 2) * 128 + i];
            a3 += dlogits[(size_t)(n + 3) * 65 + j] * hp[(size_t)(n + 3) * 128 + i];
        }
        for (; n < n1; ++n) a0 += dlogits[(size_t)n * 65 + j] * hp[(size_t)n * 128 + i];
        atomicAdd(&dlw[idx], (a0 + a1) + (a2 + a3));
    } else if (idx < n_lw + 65) {
        const int j = idx - n_lw;
        float acc = 0.f;
        for (int n = n0; n < n1; ++n) acc += dlogits[(size_t)n * 65 + j];
        atomicAdd(&dlb[j], acc);
    } else if (idx < n_lw + 65 + n_w1) {
        const int r = idx - n_lw - 65, k = r / 64, i = r % 64;
        const float wk = w2[k];
        float a0 = 0.f, a1 = 0.f;
        int n = n0;
        for (; n + 1 < n1; n += 2) {
            a0 += (u[(size_t)n * Hv + k] > 0.f ? dvpre[n] * wk : 0.f) * hv[(size_t)n * 64 + i];
            a1 += (u[(size_t)(n + 1) * Hv + k] > 0.f ? dvpre[n + 1] * wk : 0.f) * hv[(size_t)(n + 1) * 64 + i];
        }
        for (; n < n1; ++n) a0 += (u[(size_t)n * Hv + k] > 0.f ? dvpre[n] * wk : 0.f) * hv[(size_t)n * 64 + i];
        atomicAdd(&dw1[r], a0 + a1);
    } else if (idx < n_lw + 65 + n_w1 + Hv) {
        const int k = idx - n_lw - 65 - n_w1;
        const float wk = w2[k];
        float a1 = 0.f, a2 = 0.f;
        for (int n = n0; n < n1; ++n) { const float uu = u[(size_t)n * Hv + k]; a1 += uu > 0.f ? dvpre[n] * wk : 0.f; a2 += dvpre[n] * uu; }
        atomicAdd(&db1[k], a1); atomicAdd(&dw2[k], a2);
    } else if (idx == n_lw + 65 + n_w1 + Hv) {
        float acc = 0.f;
        for (int n = n0; n < n1; ++n) acc += dvpre[n];
        atomicAdd(&db2[0], acc);
    }
}

// Gradients wrt the normalised head features, masked by their ReLU: dhn[k][p].
__global__ void head_fc_bwd_input_kernel(const float *__restrict__ dlogits, const float *__restrict__ lw, const float *__restrict__ u,
                                         const float *__restrict__ dvpre, const float *__restrict__ w1, const float *__restrict__ w2,
                                         const float *__restrict__ hn, size_t P, int Hv, float *__restrict__ dhn) {
    OZ_PDL_PROLOGUE();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 3 * P) return;
    const int k = (int)(idx / P);
    const size_t p = idx % P, n = p >> 6;
    const int pos = (int)(p & 63);
    float acc = 0.f;
    if (k < 2) {
        const int i = k * 64 + pos;
        for (int j = 0; j < 65; ++j) acc += dlogits[n * 65 + j] * lw[j * 128 + i];
    } else {
        const float dv = dvpre[n];
        for (int q = 0; q < Hv; ++q) acc += (u[n * Hv + q] > 0.f ? dv * w2[q] : 0.f) * w1[q * 64 + pos];
    }
    dhn[idx] = hn[idx] > 0.f ? acc : 0.f;
}

// BatchNorm backward of the three head channels (block k): dh[k][p] (gradient wrt the 1x1 conv output), dgamma, dbeta.
__global__ void __launch_bounds__(256) head_bn_bwd_stats_kernel(const float *__restrict__ h, const float *__restrict__ dhn, size_t P,
                                                                const float *__restrict__ stat, float *__restrict__ acc) {
    OZ_PDL_PROLOGUE();
    __shared__ float scratch[32];
    const int k = blockIdx.y;
    const float mean = stat[k], inv = stat[3 + k];
    const float *hk = h + (size_t)k * P, *dk = dhn + (size_t)k * P;
    float s1 = 0.f, s2 = 0.f;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) { const float d = dk[p]; s1 += d; s2 += d * (hk[p] - mean) * inv; }
    s1 = block_sum(s1, scratch); s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) { atomicAdd(&acc[2 * k], s1); atomicAdd(&acc[2 * k + 1], s2); }
}

__global__ void __launch_bounds__(256) head_bn_bwd_kernel(const float *__restrict__ h, const float *__restrict__ dhn, size_t P, const float *__restrict__ stat,
                                   const float *__restrict__ acc, const float *__restrict__ gamma_p, const float *__restrict__ gamma_v, float *__restrict__ dh,
                                   float *__restrict__ dgamma_p, float *__restrict__ dbeta_p, float *__restrict__ dgamma_v, float *__restrict__ dbeta_v) {
    OZ_PDL_PROLOGUE();
    const int k = blockIdx.y;
    const float mean = stat[k], inv = stat[3 + k], s1 = acc[2 * k], s2 = acc[2 * k + 1];
    const float *hk = h + (size_t)k * P, *dk = dhn + (size_t)k * P;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (k < 2) { dgamma_p[k] = s2; dbeta_p[k] = s1; } else { dgamma_v[0] = s2; dbeta_v[0] = s1; }
    }
    const float g = k < 2 ? gamma_p[k] : gamma_v[0], invM = 1.f / (float)P;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x)
        dh[(size_t)k * P + p] = g * inv * (dk[p] - s1 * invM - (hk[p] - mean) * inv * s2 * invM);
}

// 1x1 head convolutions backward.  da[p][c] = sum_k dh[k][p] * w[k][c] (bf16, the gradient entering the tower);
// dw[k][c] += sum_p dh[k][p] * a[p][c]; db[k] += sum_p dh[k][p].  One block per 64 positions, one thread per channel.
__global__ void head_conv_bwd_kernel(const bf16 *__restrict__ a, const float *__restrict__ dh, size_t P, int C, const float *__restrict__ pw,
                                     const float *__restrict__ vw, bf16 *__restrict__ da, float *__restrict__ dpw, float *__restrict__ dpb,
                                     float *__restrict__ dvw, float *__restrict__ dvb) {
    OZ_PDL_PROLOGUE();
    __shared__ float sd[3][64];
    const size_t p0 = (size_t)blockIdx.x * 64;
    for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) sd[i / 64][i % 64] = p0 + (i % 64) < P ? dh[(size_t)(i / 64) * P + p0 + (i % 64)] : 0.f;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float w0 = pw[c], w1 = pw[C + c], w2 = vw[c];
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        for (int r = 0; r < 64 && p0 + r < P; ++r) {
            const float av = __bfloat162float(a[(p0 + r) * C + c]);
            g0 += sd[0][r] * av; g1 += sd[1][r] * av; g2 += sd[2][r] * av;
            da[(p0 + r) * C + c] = __float2bfloat16(sd[0][r] * w0 + sd[1][r] * w1 + sd[2][r] * w2);
        }
        atomicAdd(&dpw[c], g0); atomicAdd(&dpw[C + c], g1); atomicAdd(&dvw[c], g2);
    }
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int r = 0; r < 64; ++r) s += sd[threadIdx.x][r];
        atomicAdd(threadIdx.x < 2 ? &dpb[threadIdx.x] : &dvb[0], s);
    }
}

// ------------------------------------------------------------------ L2 term + SGD with momentum (torch.optim.SGD, dampening 0)
// g += 2 * l2 * p; loss[2] += l2 * sum p^2; if (apply) { buf = momentum * buf + g; p -= lr * buf; }
__global__ void l2_sgd_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ buf, size_t n, float l2, float lr, float momentum,
                              int apply, float *__restrict__ loss) {
    OZ_PDL_PROLOGUE();
    __shared__ float scratch[32];
    float sq = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float pv = p[i];
        sq += pv * pv;
        const float gv = g[i] + 2.f * l2 * pv;
        g[i] = gv;
        if (apply) { const float b = momentum * buf[i] + gv; buf[i] = b; p[i] = pv - lr * b; }
    }
    sq = block_sum(sq, scratch);
    if (threadIdx.x == 0) atomicAdd(&loss[2], l2 * sq);
}

unsigned grid_for(size_t items, int block) { return (unsigned)((items + block - 1) / block); }

}  // namespace
}  // namespace oz

using namespace oz;

// ------------------------------------------------------------------ trainer object
struct ConvLayerT {
    size_t w, b, gamma, beta;     // offsets (floats) into the parameter vector
    size_t rm, rv;                // offsets into the buffer vector
    int cin, cinp;
    bf16 *wf = nullptr, *wd = nullptr;
};

struct oz_trainer {
    int device = 0;
    oz_resnet_shape shape{};
    int max_batch = 0;
    size_t n_params = 0, n_buffers = 0;
    std::vector<ConvLayerT> conv;
    // head offsets (params): policy conv w,b, bn gamma,beta, linear w,b; value conv w,b, bn g,b, linear1 w,b, linear2 w,b
    size_t pcw, pcb, pg, pb, plw, plb, vcw, vcb, vg, vb, v1w, v1b, v2w, v2b;
    size_t prm, vrm;              // buffers: policy running mean[2] var[2]; value running mean[1] var[1]
    float *d_params = nullptr, *d_grads = nullptr, *d_mom = nullptr, *d_buffers = nullptr;
    float *d_features = nullptr, *d_tpolicy = nullptr, *d_tvalue = nullptr;
    bf16 *d_x0 = nullptr;
    std::vector<bf16 *> A, Z;     // per conv layer: post-activation, pre-BN
    float *d_mean = nullptr, *d_invstd = nullptr;   // [layers][C]
    float *d_sums = nullptr;      // [2C] scratch
    float *d_zero_bias = nullptr;
    PackDesc *d_pack = nullptr;
    bf16 *d_dA = nullptr, *d_dA2 = nullptr, *d_dY = nullptr, *d_dZ = nullptr;
    float *d_G = nullptr;
    float *d_h = nullptr, *d_hn = nullptr, *d_dhn = nullptr, *d_dh = nullptr, *d_stat = nullptr, *d_hp = nullptr, *d_hv = nullptr;
    float *d_policy = nullptr, *d_value = nullptr, *d_dlogits = nullptr, *d_u = nullptr, *d_dvpre = nullptr, *d_loss = nullptr;
    // CUDA-graph replay of the whole step: at batch 256 the step is some 250 launches of a few microseconds each, so the
    // host, not the GPU, sets the pace when every kernel is launched eagerly.  One graph per (batch, hyper-parameters).
    struct StepKey { int n, apply; oz_train_config cfg; };
    StepKey last_key{};
    bool have_last_key = false;
    cudaGraphExec_t graph = nullptr;
    StepKey graph_key{};
    int use_graphs = 1;
    unsigned long long graph_launches = 0;
    cudaStream_t stream = nullptr;
    std::vector<void *> allocs;
};

namespace {

template <typename T>
int talloc(oz_trainer *t, T **p, size_t n) {
    *p = nullptr;
    OZ_CUDA(cudaMalloc((void **)p, std::max<size_t>(1, n) * sizeof(T)));
    OZ_CUDA(cudaMemset(*p, 0, std::max<size_t>(1, n) * sizeof(T)));
    t->allocs.push_back((void *)*p);
    return OZ_OK;
}

void layout_params(oz_trainer *t) {
    const size_t C = t->shape.channels, Hv = t->shape.value_hidden;
    const int L = 1 + 2 * t->shape.blocks;
    size_t o = 0, bo = 0;
    t->conv.resize(L);
    for (int l = 0; l < L; ++l) {
        ConvLayerT &c = t->conv[l];
        c.cin = l == 0 ? t->shape.in_planes : (int)C;
        c.cinp = l == 0 ? 32 : (int)C;
        c.w = o; o += C * c.cin * 9;
        c.b = o; o += C;
        c.gamma = o; o += C;
        c.beta = o; o += C;
        c.rm = bo; bo += C;
        c.rv = bo; bo += C;
    }
    t->pcw = o; o += 2 * C; t->pcb = o; o += 2; t->pg = o; o += 2; t->pb = o; o += 2; t->plw = o; o += 65 * 128; t->plb = o; o += 65;
    t->vcw = o; o += C; t->vcb = o; o += 1; t->vg = o; o += 1; t->vb = o; o += 1; t->v1w = o; o += Hv * 64; t->v1b = o; o += Hv;
    t->v2w = o; o += Hv; t->v2b = o; o += 1;
    t->prm = bo; bo += 4; t->vrm = bo; bo += 2;
    t->n_params = o; t->n_buffers = bo;
}

int pack_all_weights(oz_trainer *t) {
    const int C = t->shape.channels, L = (int)t->conv.size();
    launch_pdl(pack_weights_kernel, dim3(dim3(grid_for((size_t)9 * C * C, 256), L)), dim3(256), 0, t->stream, t->d_params, t->d_pack, C);
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

// One training step on the inputs staged in the trainer's own buffers (features float [n][in_planes][64], policy targets
// [n][65], value targets [n]).  Only stream work: this is what the step graph captures.
int run_step(oz_trainer *t, int n, const oz_train_config *cfg, int apply) {
    const float *d_features = t->d_features, *d_tp = t->d_tpolicy, *d_tv = t->d_tvalue;
    const int C = t->shape.channels, Hv = t->shape.value_hidden, L = (int)t->conv.size();
    const size_t P = (size_t)n * 64;
    cudaStream_t s = t->stream;
    float *par = t->d_params, *grd = t->d_grads, *buf = t->d_buffers;
    OZ_CUDA(cudaMemsetAsync(grd, 0, t->n_params * sizeof(float), s));
    OZ_CUDA(cudaMemsetAsync(t->d_loss, 0, 16 * sizeof(float), s));      // 4 losses + the head BatchNorm accumulators
    OZ_CUDA(cudaMemsetAsync(t->d_sums, 0, (size_t)2 * L * 2 * C * kBnCopies * sizeof(float), s));     // reduction slots of every layer, forward and backward
    OZ_TRY(pack_all_weights(t));
    launch_planes_to_nhwc32(d_features, t->shape.in_planes, (size_t)n, (size_t)n, t->d_x0, s);
    const int bn_block = 256, bn_rows = 256 / (C / 8);
    const size_t bn_chunks = (P + (size_t)bn_rows * kBnUnroll - 1) / ((size_t)bn_rows * kBnUnroll);   // every thread gets kBnUnroll rows or more
    const int bn_grid = (int)std::min<size_t>(1184, bn_chunks);       // streaming passes: up to 8 blocks per SM
    const int bn_rgrid = (int)std::min<size_t>(592, bn_chunks);       // reductions: fewer blocks, fewer atomics (37 per address and copy)

    // ---- forward through the tower (neural_net.py:9-65, train mode)
    for (int l = 0; l < L; ++l) {
        ConvLayerT &c = t->conv[l];
        const bf16 *in = l == 0 ? t->d_x0 : t->A[l - 1];
        // the convolution's epilogue also accumulates the BatchNorm batch statistics of its fp32 results (STATS instantiation)
        float *fsum = t->d_sums + (size_t)l * 2 * C * kBnCopies;
        OZ_TRY(ResNet::conv3x3(in, c.cinp, c.wf, t->d_zero_bias, nullptr, 0, t->Z[l], C, (size_t)n, s, nullptr, fsum, kBnCopies));
        const bf16 *skip = (l >= 2 && l % 2 == 0) ? t->A[l - 2] : nullptr;     // second conv of a residual block
        launch_pdl(bn_apply_kernel, dim3(bn_grid), dim3(bn_block), 0, s, t->Z[l], skip, P, C, fsum, par + c.gamma, par + c.beta, par + c.b, cfg->bn_eps,
                                                            cfg->bn_momentum, t->d_mean + (size_t)l * C, t->d_invstd + (size_t)l * C, buf + c.rm,
                                                            buf + c.rv, t->A[l]);
    }
    // ---- heads forward + losses (neural_net.py:68-128, train.py:494-495)
    const bf16 *a_last = t->A[L - 1];
    launch_pdl(head_conv_fwd_kernel, dim3(grid_for(P * 32, 256)), dim3(256), 0, s, a_last, P, C, par + t->pcw, par + t->pcb, par + t->vcw, par + t->vcb, t->d_h);
    const dim3 hbn_grid((unsigned)std::max<size_t>(1, std::min<size_t>(64, P / 2048)), 3);
    float *hacc_f = t->d_loss + 4, *hacc_b = t->d_loss + 10;       // [3][2] accumulators, zeroed with the losses
    launch_pdl(head_bn_stats_kernel, hbn_grid, dim3(256), 0, s, t->d_h, P, hacc_f);
    launch_pdl(head_bn_fwd_kernel, hbn_grid, dim3(256), 0, s, t->d_h, P, hacc_f, par + t->pg, par + t->pb, par + t->vg, par + t->vb, cfg->bn_eps, cfg->bn_momentum, buf + t->prm,
                                          buf + t->vrm, t->d_stat, t->d_hn, t->d_hp, t->d_hv);
    const size_t fc_smem = (size_t)(65 * kFcLwStride + Hv * kFcW1Stride + 8 * 192) * sizeof(float);
    launch_pdl(head_fc_fwd_kernel, dim3((unsigned)std::min<size_t>(148, (n + 7) / 8)), dim3(256), fc_smem, s, t->d_hp, t->d_hv, n, Hv, par + t->plw, par + t->plb, par + t->v1w, par + t->v1b,
                                                                     par + t->v2w, par + t->v2b, d_tp, d_tv, t->d_policy, t->d_value, t->d_dlogits,
                                                                     t->d_u, t->d_dvpre, t->d_loss);
    // ---- heads backward
    const int n_head_params = 65 * 128 + 65 + Hv * 64 + Hv + 1;
    launch_pdl(head_fc_bwd_params_kernel, dim3(grid_for(n_head_params, 128), (unsigned)std::max(1, std::min(32, n / 64))), dim3(128), 0, s, t->d_hp, t->d_hv, t->d_dlogits, t->d_u, t->d_dvpre, par + t->v2w, n, Hv,
                                                                           grd + t->plw, grd + t->plb, grd + t->v1w, grd + t->v1b, grd + t->v2w,
                                                                           grd + t->v2b);
    launch_pdl(head_fc_bwd_input_kernel, dim3(grid_for(3 * P, 256)), dim3(256), 0, s, t->d_dlogits, par + t->plw, t->d_u, t->d_dvpre, par + t->v1w, par + t->v2w, t->d_hn, P,
                                                                  Hv, t->d_dhn);
    launch_pdl(head_bn_bwd_stats_kernel, hbn_grid, dim3(256), 0, s, t->d_h, t->d_dhn, P, t->d_stat, hacc_b);
    launch_pdl(head_bn_bwd_kernel, hbn_grid, dim3(256), 0, s, t->d_h, t->d_dhn, P, t->d_stat, hacc_b, par + t->pg, par + t->vg, t->d_dh, grd + t->pg, grd + t->pb, grd + t->vg,
                                          grd + t->vb);
    bf16 *dA = t->d_dA, *dA_other = t->d_dA2;
    launch_pdl(head_conv_bwd_kernel, dim3(grid_for(P, 64)), dim3(128), 0, s, a_last, t->d_dh, P, C, par + t->pcw, par + t->vcw, dA, grd + t->pcw, grd + t->pcb, grd + t->vcw,
                                                         grd + t->vcb);
    // ---- tower backward
    for (int l = L - 1; l >= 0; --l) {
        ConvLayerT &c = t->conv[l];
        const float *mean = t->d_mean + (size_t)l * C, *inv = t->d_invstd + (size_t)l * C;
        const bool block_out = l >= 2 && l % 2 == 0;            // output of a residual block: dy also flows into the skip connection
        float *bsum = t->d_sums + (size_t)(L + l) * 2 * C * kBnCopies;
        launch_pdl(bn_bwd_reduce_kernel, dim3(bn_rgrid), dim3(bn_block), 0, s, dA, t->A[l], t->Z[l], P, C, mean, inv, bsum);
        launch_pdl(bn_bwd_apply_kernel, dim3(bn_grid), dim3(bn_block), 0, s, dA, t->A[l], t->Z[l], P, C, mean, inv, par + c.gamma, bsum, t->d_dZ,
                                                                block_out ? t->d_dY : nullptr, grd + c.gamma, grd + c.beta);
        // weight gradient: dW[co][tap][ci] = sum_p dZ[p][co] * cols[p][tap][ci]
        const bf16 *in = l == 0 ? t->d_x0 : t->A[l - 1];
        OZ_TRY(launch_wgrad3x3(in, c.cinp, t->d_dZ, C, (size_t)n, t->d_G, s));
        launch_pdl(wgrad_scatter_kernel, dim3(grid_for((size_t)C * c.cin * 9, 256)), dim3(256), 0, s, t->d_G, C, c.cin, c.cinp, grd + c.w);
        // (the conv bias has no gradient besides the L2 term: BatchNorm removes it)
        if (l == 0) break;
        // input gradient: the same convolution with rotated taps; the second-to-last layer of a block adds the skip gradient
        const bool block_in = l % 2 == 1;                       // first conv of a residual block: its input is the block input
        OZ_TRY(ResNet::conv3x3(t->d_dZ, C, c.wd, t->d_zero_bias, block_in ? t->d_dY : nullptr, 0, dA_other, C, (size_t)n, s));
        std::swap(dA, dA_other);
    }
    // ---- L2 term, optimizer
    launch_pdl(l2_sgd_kernel, dim3(296), dim3(256), 0, s, par, grd, t->d_mom, t->n_params, cfg->l2_weight, cfg->lr, cfg->momentum, apply, t->d_loss);
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

bool same_key(const oz_trainer::StepKey &a, const oz_trainer::StepKey &b) {
    return a.n == b.n && a.apply == b.apply && std::memcmp(&a.cfg, &b.cfg, sizeof(oz_train_config)) == 0;
}

void drop_step_graph(oz_trainer *t) {
    if (t->graph) { cudaGraphExecDestroy(t->graph); t->graph = nullptr; }
    t->have_last_key = false;
}

// Inputs are staged; run the step eagerly the first time a (batch, hyper-parameter) combination is seen, capture it the
// second time, replay it from then on.
int execute_step(oz_trainer *t, int n, const oz_train_config *cfg, int apply, float *losses_host) {
    if (n < 1 || n > t->max_batch || n % kBoardPad != 0)
        return fail(OZ_ERR_INVALID_ARGUMENT, "Expected a batch size in [32, " + std::to_string(t->max_batch) + "] that is a multiple of 32, but got " + std::to_string(n) + ".");
    oz_trainer::StepKey key{};
    key.n = n; key.apply = apply; key.cfg = *cfg;
    cudaStream_t s = t->stream;
    if (t->use_graphs && t->graph && same_key(key, t->graph_key)) {
        OZ_CUDA(cudaGraphLaunch(t->graph, s));
        count_launch((int)t->graph_launches);
    } else if (t->use_graphs && t->have_last_key && same_key(key, t->last_key)) {
        if (t->graph) { cudaGraphExecDestroy(t->graph); t->graph = nullptr; }
        cudaGraph_t g = nullptr;
        const unsigned long long before = g_kernel_launches;
        OZ_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const int st = run_step(t, n, cfg, apply);
        const cudaError_t ce = cudaStreamEndCapture(s, &g);
        t->graph_launches = g_kernel_launches - before;
        g_kernel_launches = before;                          // capturing launched nothing
        if (st || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            return st ? st : fail(OZ_ERR_CUDA, std::string("graph capture of the training step failed: ") + cudaGetErrorString(ce));
        }
        const cudaError_t ie = cudaGraphInstantiate(&t->graph, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) { t->graph = nullptr; return fail(OZ_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
        t->graph_key = key;
        OZ_CUDA(cudaGraphLaunch(t->graph, s));
        count_launch((int)t->graph_launches);
    } else {
        OZ_TRY(run_step(t, n, cfg, apply));
        t->last_key = key; t->have_last_key = true;
    }
    if (losses_host) {
        float h[4];
        OZ_CUDA(cudaMemcpyAsync(h, t->d_loss, sizeof(h), cudaMemcpyDeviceToHost, s));
        OZ_CUDA(cudaStreamSynchronize(s));
        losses_host[0] = h[0] + h[1] + h[2]; losses_host[1] = h[0]; losses_host[2] = h[1]; losses_host[3] = h[2];   // total, policy, value, l2
    }
    return OZ_OK;
}

}  // namespace

extern "C" {

void oz_train_config_default(oz_train_config *c) {
    c->lr = 0.001f; c->momentum = 0.9f; c->l2_weight = 1e-4f;          // train.py:103-112, 202-206
    c->bn_momentum = 0.1f; c->bn_eps = 1e-5f;                          // nn.BatchNorm2d defaults
}

int oz_trainer_create(int device, const oz_resnet_shape *shape, int max_batch, oz_trainer **out) {
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(OZ_ERR_CUDA, "no such CUDA device: the trainer has no CPU path");
    if (shape->channels < 128 || shape->channels % 128 != 0 || shape->channels > 384)
        return fail(OZ_ERR_INVALID_ARGUMENT, "conv_channels must be 128, 256 or 384 for the native trainer");
    if (shape->in_planes < 1 || shape->in_planes > 32 || shape->blocks < 0 || shape->value_hidden < 1)
        return fail(OZ_ERR_INVALID_ARGUMENT, "unsupported network shape");
    if (max_batch < kBoardPad || max_batch % kBoardPad != 0) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected max_batch to be a positive multiple of 32.");
    OZ_CUDA(cudaSetDevice(device));
    oz_trainer *t = new oz_trainer();
    t->device = device; t->shape = *shape; t->max_batch = max_batch;
    layout_params(t);
    int st = OZ_OK;
    do {
        if (cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess) { st = fail(OZ_ERR_CUDA, "stream creation failed"); break; }
        const size_t C = shape->channels, Hv = shape->value_hidden, N = max_batch, P = N * 64;
        const int L = (int)t->conv.size();
        if ((st = talloc(t, &t->d_params, t->n_params)) || (st = talloc(t, &t->d_grads, t->n_params)) || (st = talloc(t, &t->d_mom, t->n_params)) ||
            (st = talloc(t, &t->d_buffers, t->n_buffers)) || (st = talloc(t, &t->d_features, N * shape->in_planes * 64)) ||
            (st = talloc(t, &t->d_tpolicy, N * 65)) || (st = talloc(t, &t->d_tvalue, N)) || (st = talloc(t, &t->d_x0, P * 32)) ||
            (st = talloc(t, &t->d_mean, L * C)) || (st = talloc(t, &t->d_invstd, L * C)) || (st = talloc(t, &t->d_sums, (size_t)2 * L * 2 * C * kBnCopies)) ||
            (st = talloc(t, &t->d_zero_bias, C)) || (st = talloc(t, &t->d_dA, P * C)) || (st = talloc(t, &t->d_dA2, P * C)) ||
            (st = talloc(t, &t->d_dY, P * C)) || (st = talloc(t, &t->d_dZ, P * C)) ||
            (st = talloc(t, &t->d_G, C * 9 * C)) || (st = talloc(t, &t->d_h, 3 * P)) || (st = talloc(t, &t->d_hn, 3 * P)) ||
            (st = talloc(t, &t->d_dhn, 3 * P)) || (st = talloc(t, &t->d_dh, 3 * P)) || (st = talloc(t, &t->d_stat, 8)) ||
            (st = talloc(t, &t->d_hp, N * 128)) || (st = talloc(t, &t->d_hv, N * 64)) || (st = talloc(t, &t->d_policy, N * 65)) ||
            (st = talloc(t, &t->d_value, N)) || (st = talloc(t, &t->d_dlogits, N * 65)) || (st = talloc(t, &t->d_u, N * Hv)) ||
            (st = talloc(t, &t->d_dvpre, N)) || (st = talloc(t, &t->d_loss, 16)))
            break;
        if (cudaFuncSetAttribute(head_fc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((65 * kFcLwStride + Hv * kFcW1Stride + 8 * 192) * sizeof(float))) != cudaSuccess) {
            st = fail(OZ_ERR_CUDA, "value_head_hidden_channels too large for the trainer's head kernel"); break;
        }
        t->A.resize(L); t->Z.resize(L);
        for (int l = 0; l < L && !st; ++l) {
            ConvLayerT &c = t->conv[l];
            if ((st = talloc(t, &t->A[l], P * C)) || (st = talloc(t, &t->Z[l], P * C)) || (st = talloc(t, &c.wf, (size_t)9 * C * c.cinp))) break;
            if (l > 0 && (st = talloc(t, &c.wd, (size_t)9 * C * C))) break;
        }
        if (st) break;
        std::vector<PackDesc> desc(L);
        for (int l = 0; l < L; ++l) desc[l] = PackDesc{t->conv[l].w, t->conv[l].cin, t->conv[l].cinp, t->conv[l].wf, t->conv[l].wd};
        if ((st = talloc(t, &t->d_pack, (size_t)L))) break;
        if (cudaMemcpy(t->d_pack, desc.data(), L * sizeof(PackDesc), cudaMemcpyHostToDevice) != cudaSuccess) { st = fail(OZ_ERR_CUDA, "descriptor upload failed"); break; }
    } while (0);
    if (st) { oz_trainer_destroy(t); return st; }
    *out = t;
    return OZ_OK;
}

void oz_trainer_destroy(oz_trainer *t) {
    if (!t) return;
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamSynchronize(t->stream);
    if (t->graph) cudaGraphExecDestroy(t->graph);
    for (void *p : t->allocs) cudaFree(p);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

size_t oz_trainer_num_params(const oz_trainer *t) { return t->n_params; }
size_t oz_trainer_num_buffers(const oz_trainer *t) { return t->n_buffers; }

int oz_trainer_set_state(oz_trainer *t, const float *params, const float *buffers, const float *momentum) {
    OZ_CUDA(cudaSetDevice(t->device));
    OZ_CUDA(cudaStreamSynchronize(t->stream));
    if (params) OZ_CUDA(cudaMemcpy(t->d_params, params, t->n_params * sizeof(float), cudaMemcpyHostToDevice));
    if (buffers) OZ_CUDA(cudaMemcpy(t->d_buffers, buffers, t->n_buffers * sizeof(float), cudaMemcpyHostToDevice));
    if (momentum) OZ_CUDA(cudaMemcpy(t->d_mom, momentum, t->n_params * sizeof(float), cudaMemcpyHostToDevice));
    return OZ_OK;
}

int oz_trainer_get_state(oz_trainer *t, float *params, float *buffers, float *momentum, float *grads) {
    OZ_CUDA(cudaSetDevice(t->device));
    OZ_CUDA(cudaStreamSynchronize(t->stream));
    if (params) OZ_CUDA(cudaMemcpy(params, t->d_params, t->n_params * sizeof(float), cudaMemcpyDeviceToHost));
    if (buffers) OZ_CUDA(cudaMemcpy(buffers, t->d_buffers, t->n_buffers * sizeof(float), cudaMemcpyDeviceToHost));
    if (momentum) OZ_CUDA(cudaMemcpy(momentum, t->d_mom, t->n_params * sizeof(float), cudaMemcpyDeviceToHost));
    if (grads) OZ_CUDA(cudaMemcpy(grads, t->d_grads, t->n_params * sizeof(float), cudaMemcpyDeviceToHost));
    return OZ_OK;
}

int oz_trainer_step(oz_trainer *t, const float *features, const float *policy, const float *value, int n, const oz_train_config *cfg,
                    int apply_update, float *losses) {
    OZ_CUDA(cudaSetDevice(t->device));
    if (n < 1 || n > t->max_batch) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected 1 <= batch <= max_batch, but got " + std::to_string(n) + ".");
    OZ_CUDA(cudaMemcpyAsync(t->d_features, features, (size_t)n * t->shape.in_planes * 64 * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    OZ_CUDA(cudaMemcpyAsync(t->d_tpolicy, policy, (size_t)n * 65 * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    OZ_CUDA(cudaMemcpyAsync(t->d_tvalue, value, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    return execute_step(t, n, cfg, apply_update, losses);
}

int oz_trainer_step_device(oz_trainer *t, const float *features, const float *policy, const float *value, int n, const oz_train_config *cfg,
                           int apply_update, float *losses) {
    NvtxRange nvtx_step("oz: training step");
    OZ_CUDA(cudaSetDevice(t->device));
    if (n < 1 || n > t->max_batch) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected 1 <= batch <= max_batch, but got " + std::to_string(n) + ".");
    // staged into the trainer's own buffers so that the captured step always reads the same addresses
    OZ_CUDA(cudaMemcpyAsync(t->d_features, features, (size_t)n * t->shape.in_planes * 64 * sizeof(float), cudaMemcpyDeviceToDevice, t->stream));
    OZ_CUDA(cudaMemcpyAsync(t->d_tpolicy, policy, (size_t)n * 65 * sizeof(float), cudaMemcpyDeviceToDevice, t->stream));
    OZ_CUDA(cudaMemcpyAsync(t->d_tvalue, value, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, t->stream));
    return execute_step(t, n, cfg, apply_update, losses);
}

int oz_trainer_set_graphs(oz_trainer *t, int on) {
    t->use_graphs = on != 0;
    if (!on) drop_step_graph(t);
    return OZ_OK;
}

int oz_trainer_outputs(oz_trainer *t, int n, float *policy, float *value) {
    OZ_CUDA(cudaSetDevice(t->device));
    OZ_CUDA(cudaStreamSynchronize(t->stream));
    if (n < 1 || n > t->max_batch) return fail(OZ_ERR_INVALID_ARGUMENT, "batch out of range");
    if (policy) OZ_CUDA(cudaMemcpy(policy, t->d_policy, (size_t)n * 65 * sizeof(float), cudaMemcpyDeviceToHost));
    if (value) OZ_CUDA(cudaMemcpy(value, t->d_value, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return OZ_OK;
}

void *oz_trainer_stream(oz_trainer *t) { return (void *)t->stream; }

}  // extern "C"
