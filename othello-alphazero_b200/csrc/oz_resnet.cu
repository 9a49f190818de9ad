// ResNet policy/value evaluator for sm_100a (reference: neural_net.py:9-172, eval mode).
//
// conv3x3_pair_kernel: implicit-GEMM 3x3 "same" convolution, bf16 in / fp32 accumulate, issued for a CTA PAIR
// (tcgen05.mma cta_group::2, M = 256, N = 128, K = 16; accumulators in TMEM, operands staged by TMA); see the
// comment above the kernel.  heads_kernel: 1x1 convs + FC + softmax / tanh for both heads, one warp per board.
#include <cuda_bf16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "oz_resnet.cuh"
#include "oz_tc.cuh"

namespace oz {

// ------------------------------------------------------------------ geometry
constexpr int kBoardsPerUnit = 16;
constexpr int kKBlock = 32;                          // channels per K block (64 B rows)
constexpr int kBoxX = 9;                             // halo: x = 8 is the zero column
constexpr int kRowBytes = kKBlock * 2;               // 64
constexpr int kSlotBytes = kBoardsPerUnit * kRowBytes;            // 1024: one (y, x) slot
constexpr int kTmemCols = 512;

// K-major, SWIZZLE_64B shared-memory matrix descriptor (rows of 64 B, 8-row atoms of 512 B).
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;            // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                     // SWIZZLE_64B
    return d;
}
struct ConvParams {
    const __nv_bfloat16 *skip;   // NHWC, may be null
    __nv_bfloat16 *out;          // NHWC [boards][8][8][cout_total]
    const float *bias;           // [cout_total]
    int units;                   // boards / 16
    int kblocks;                 // cin / 32
    int cout_total;              // multiple of 128
    int relu;
    const float *head_w;         // pair kernel, last layer of a 128-channel tower: [3][128] 1x1 head-conv weights (policy 0, policy 1, value)
    float *head_out;             // ... and their pre-activations [boards][64][4] (zeroed by the caller; two column halves add up)
    int *work;                   // dynamic pass scheduler: this launch's pass counter (zeroed by the caller), or null = static round-robin
    float *stats;                // STATS instantiation (training forward): [copies][2 * cout_total] sum / sum of squares of the fp32 outputs
    int stats_copies;            // ... accumulated into copy blockIdx.x % stats_copies
#ifdef OZ_CONV_TIMELINE
    unsigned long long *tl;      // measurement build only: [8] %globaltimer stamps of CTA 0 for this launch
#endif
};
#ifdef OZ_CONV_TIMELINE
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define OZ_TL(k, cond) do { if (p.tl && blockIdx.x == 0 && (cond)) p.tl[k] = gtime(); } while (0)
#else
#define OZ_TL(k, cond) do { } while (0)
#endif

// ------------------------------------------------------------------ CTA-pair kernel (cta_group::2)
// conv3x3_pair_kernel: the same implicit GEMM issued for a CTA PAIR (cluster of 2, one TPC):
//   * D[M = 256 (128 rows per CTA), N = 128 couts] += A * B^T with tcgen05.mma.cta_group::2.  Each CTA
//     stages its own 128-row activation tiles and only HALF of the weight tile (64 couts); the pair's
//     tensor cores read B from both shared memories, so per-SM shared-memory traffic per MMA drops from
//     8 KB (128 B/clk, the cta_group::1 ceiling) to 6 KB.
//   * Weights stay resident: for C = 128 a CTA's half of a layer is 9 taps x 4 K blocks x 4 KB = 144 KB,
//     loaded once per launch.  Wider towers stream the same 4 KB tiles through the 36 slots as a ring.
//   * A pass is 16 boards x 2 board rows (2 accumulator tiles of 128 columns); TMEM holds two passes,
//     so the epilogue of pass p overlaps the MMAs of pass p + 1.
//   * Eight epilogue warps (TMEM lane quarter x column half); the residual input is prefetched into
//     registers before the accumulators are ready.
namespace pair {
constexpr int kYPerPass = 2;
constexpr int kBoxYp = kYPerPass + 2;
constexpr int kActBoxBytesP = kBoxX * kBoxYp * kSlotBytes;          // 36864
// Stage layout: [guard slot][box 0][box 1].  The guard slot (zeros) is what tap (-1, -1) reads in front of box 0; in front
// of box 1 sits the last slot of box 0 = (last row, x = 8), the out-of-bounds column TMA always fills with zeros.
constexpr int kActStagesP = 2;
constexpr int kWTileBytes = 64 * kRowBytes;                         // 64 couts x 32 cin = 4096
constexpr int kWSlots = 36;
constexpr int kWGroups = kWSlots / 9;                                 // the weight ring is handed over in groups of nine tap tiles (one K block)
constexpr int kEpiWarps = 8;
constexpr int kThreadsP = (3 + kEpiWarps) * 32;
constexpr int kMiscBytes = 3072;                                    // barriers + scheduler ring (1 KB), bias / head-conv weights (2 KB)
// Ring of published pass indices (dynamic scheduler; see the kernel).
constexpr int kSchedSlots = 8;
constexpr int kSmemBytesP = 1024 + kSlotBytes + kActStagesP * kActBoxBytesP + kWSlots * kWTileBytes + kMiscBytes;
// 233 472 B per SM, 1 KB of it reserved per resident CTA: what this kernel leaves free (6 144 B) must hold one block of the
// narrow select kernel (5 120 B static + its 1 KB), or the pipelined search cannot run tree kernels under the convolutions.
static_assert(kSmemBytesP + 1024 + 5120 + 1024 <= 233472, "shared memory budget (co-residency with select_kernel<1>)");
// kind::f16 instruction descriptor: bf16 x bf16 -> f32, K-major operands, M = 256 (pair), N = 128.
constexpr uint32_t kIdescP = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
}  // namespace pair

// The same wait with cluster-scope acquire: the data the barrier guards was written by the peer CTA (st.shared::cluster).
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((spins & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000ll) { printf("conv3x3_pair_kernel: scheduler barrier timeout (block %d thread %d bar %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar); __trap(); }
        }
    }
}
__device__ __forceinline__ void st_cluster_b32(uint32_t cluster_addr, int a) {
    asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(cluster_addr), "r"(a) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 256-bit global accesses (sm_100): one full 32-byte sector per lane per instruction.
__device__ __forceinline__ void ldg256(const void *ptr, uint32_t *v) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(ptr) : "memory");
}
__device__ __forceinline__ void stg256(void *ptr, const uint32_t *v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on the same barrier offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// YPP = board rows per pass (and accumulator tiles per pass).  2 is the throughput configuration described above; 1 halves
// the serial MMA chain of a pass and doubles the number of passes, for batches that cannot give every CTA pair a pass
// otherwise (the single-game MCTS of the reference's CLIs: 32 boards are 4 pair-passes at YPP = 2).
// STATS (training forward pass): the epilogue also accumulates, per output channel, the sum and the sum of squares of the
// fp32 results (BatchNorm batch statistics, nn.BatchNorm2d in train mode) -- a 32 x 32 transpose-reduce across the lanes of a
// warp (31 shuffles per 32 channels: at step s a lane keeps half of its values and receives its partner's copies of them;
// after five steps lane j holds channel j's total over the 32 rows), kept in registers over all passes of the CTA and added
// to the global accumulators once at the end.  Replaces a separate pass over the bf16 tensor (bn_stats_kernel).  The
// residual input is compiled out of this instantiation (a training forward convolution has none).
template <int YPP, bool STATS = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pair::kThreadsP, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap map_act, const __grid_constant__ CUtensorMap map_w, const ConvParams p) {
    using namespace pair;
    constexpr int kYPerPass = YPP, kRowPasses = 8 / YPP;                      // passes per board column of rows
    constexpr int kActBoxBytesP = kBoxX * (YPP + 2) * kSlotBytes;            // the box carries one halo row above and below
    static_assert(kActBoxBytesP <= pair::kActBoxBytesP && kActBoxBytesP % 1024 == 0, "box size");
    extern __shared__ uint8_t smem_raw[];
    OZ_TL(0, threadIdx.x == 0);
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t act0 = base;
    const uint32_t w0 = act0 + kSlotBytes + kActStagesP * pair::kActBoxBytesP;
    const uint32_t misc = w0 + kWSlots * kWTileBytes;
    const uint32_t bar_act_full = misc, bar_act_empty = misc + 16;          // [2] each
    const uint32_t bar_tmem_full = misc + 32, bar_tmem_empty = misc + 48;    // [2] each
    const uint32_t bar_w_full = misc + 64, bar_w_empty = misc + 64 + 8 * kWSlots;
    uint32_t *tmem_ptr = (uint32_t *)(gen + (misc - base) + 704);
    const uint32_t bar_sched_full = misc + 720, bar_sched_empty = misc + 784;   // [kSchedSlots] each (dynamic pass scheduler)
    const uint32_t sched_ring = misc + 848;                                     // [kSchedSlots] pass indices, 8 bytes apart
    float *s_bias = (float *)(gen + (misc - base) + 1024);                   // up to 384 floats
    float *s_head = s_bias + 128;                                            // [3][128] head-conv weights (fused heads: cout_total = 128)

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int ntiles = p.cout_total >> 7;
    const int num_pp = (p.units / 2) * kRowPasses * ntiles;  // pair-passes: (unit pair, row group, cout tile)
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const bool resident = ntiles == 1 && p.kblocks <= kWGroups;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kActStagesP; ++i) { mbar_init(bar_act_full + 8 * i, 1); mbar_init(bar_act_empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_tmem_full + 8 * i, 1); mbar_init(bar_tmem_empty + 8 * i, 2 * kEpiWarps); }
        for (int i = 0; i < kWSlots; ++i) { mbar_init(bar_w_full + 8 * i, 1); mbar_init(bar_w_empty + 8 * i, 1); }
        for (int i = 0; i < kSchedSlots; ++i) { mbar_init(bar_sched_full + 8 * i, 1); mbar_init(bar_sched_empty + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_act) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kSlotBytes / 16; i += kThreadsP)                    // zero the guard slot
        *(uint4 *)(gen + i * 16) = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < p.cout_total; i += kThreadsP) s_bias[i] = p.bias[i];
    if (p.head_out) for (int i = threadIdx.x; i < 3 * 128; i += kThreadsP) s_head[i] = p.head_w[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    // Programmatic dependent launch: the next layer's CTAs may become resident as ours retire and run their own
    // set-up (barriers, TMEM, weight tiles) under our tail; they touch activations only after griddep_wait().
    griddep_launch_dependents();
    OZ_TL(1, threadIdx.x == 0);

    // Pass scheduler.  Static (p.work == nullptr): cluster c takes passes c, c + clusters, ...  Dynamic: the passes of a
    // launch are handed out in arrival order from one global counter, so a cluster that is slowed down -- a tree kernel
    // of another game group co-resident on its SMs (pipelined search), a late start behind the previous layer's tail --
    // simply takes fewer passes instead of holding the whole launch back.  The scheduler is the one warp of the pair
    // that has no other job (warp 1 of the non-leader CTA; only the leader issues MMAs): it draws an index with an
    // atomic, publishes it into a small ring in BOTH CTAs' shared memory (cluster-scope release / acquire on the ring's
    // full barriers) and stays at most two entries ahead of what the leader's activation producer reports to have read,
    // so a pair never holds more than two passes it has not started.  Every role reads the same sequence from the ring; an
    // index >= num_pp is the end marker.  None of the working roles executes an atomic, a remote store or a release on
    // this path -- the activation pipeline is two stages deep and latency-critical: with the draw and the publication
    // in the producer's loop the launch was 10 % slower, with release-arrives in the MMA thread 5 % more.
    // Ring safety (8 slots): when the producer reads entry i the MMA thread has read entry i - 3 or later (two stages
    // of run-ahead) and the epilogue entry i - 5 or later (two accumulator stages), so entry i + 2 never overwrites an
    // entry (i - 6) somebody still needs.
    const bool dynamic = p.work != nullptr;
    auto next_pass = [&](uint32_t si, int &pp) {
        if (!dynamic) { pp = cluster_id + (int)si * num_clusters; return; }
        const uint32_t slot = si % kSchedSlots;
        mbar_wait_cluster(bar_sched_full + 8 * slot, (si / kSchedSlots) & 1);
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(pp) : "r"(sched_ring + 8 * slot) : "memory");
    };
    // The pass after entry si, if it is known already (the scheduler runs ahead of the epilogue); num_pp otherwise.
    auto peek_pass = [&](uint32_t si, int pp) -> int {
        if (!dynamic) return pp + num_clusters;
        const uint32_t slot = (si + 1) % kSchedSlots;
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar_sched_full + 8 * slot), "r"(((si + 1) / kSchedSlots) & 1) : "memory");
        int v = num_pp;
        if (done) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sched_ring + 8 * slot) : "memory");
        return v;
    };

    // Roles run with warp-uniform control flow and elect one lane only around the asynchronous
    // instruction itself, so descriptors and coordinates stay in uniform registers (a divergent
    // `if (lane == 0)` loop makes the compiler re-broadcast every operand before every MMA).
    if (warp == 0) {
        // ===== activation TMA producer (one per CTA; bytes complete on the LEADER's barrier) =====
        uint32_t it = 0;
        const uint32_t full0 = mapa_rank(bar_act_full, 0);
        const uint32_t sched_go = mapa_rank(bar_sched_empty, 1);   // the scheduler's "entry read" barriers live in the peer CTA
        griddep_wait();                                   // the previous layer's output is complete and visible
        OZ_TL(2, lane == 0);
        for (uint32_t si = 0;; ++si) {
            int pp;
            next_pass(si, pp);
            if (dynamic && leader && lane == 0)           // entry si is in use: the scheduler may draw the next one
                asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(sched_go + 8 * (si % kSchedSlots)) : "memory");
            if (pp >= num_pp) break;
            const int up = pp / ntiles, unit = 2 * (up / kRowPasses) + (int)rank, yp = up % kRowPasses;
            for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                const uint32_t s = it % kActStagesP, ph = (it / kActStagesP) & 1;
                mbar_wait_b(bar_act_empty + 8 * s, ph ^ 1);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(bar_act_full + 8 * s, 2 * kActBoxBytesP);
                    tma_load_4d_pair(act0 + kSlotBytes + s * kActBoxBytesP, &map_act, full0 + 8 * s,
                                     kb * kKBlock, unit * kBoardsPerUnit, 0, yp * kYPerPass - 1);
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ===== weight TMA producer: this CTA's 64-cout half of every (K block, tap) tile =====
        // The ring is handed over in GROUPS of nine tap tiles (one K block, 36 KB; four groups): one barrier per group.
        uint32_t wgit = 0;
        const uint32_t full0 = mapa_rank(bar_w_full, 0);
        for (uint32_t si = 0;; ++si) {
            int pp;
            next_pass(si, pp);
            if (pp >= num_pp) break;
            if (resident && wgit >= (uint32_t)p.kblocks) { if (!dynamic) break; continue; }   // loaded once; keep reading the schedule
            const int nt = pp % ntiles;
            for (int kb = 0; kb < p.kblocks; ++kb, ++wgit) {
                const uint32_t g = wgit % kWGroups, ph = (wgit / kWGroups) & 1;
                mbar_wait_b(bar_w_empty + 8 * g, ph ^ 1);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(bar_w_full + 8 * g, 2 * 9 * kWTileBytes);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
                        tma_load_2d_pair(w0 + (g * 9 + tap) * kWTileBytes, &map_w, full0 + 8 * g,
                                         kb * kKBlock, tap * p.cout_total + nt * 128 + (int)rank * 64);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected thread of the leader CTA drives both tensor cores =====
        if (leader) {
            uint32_t ait = 0, wgit = 0;
            for (uint32_t np = 0;; ++np) {
                int pp;
                next_pass(np, pp);
                if (pp >= num_pp) break;
                const uint32_t acc = np & 1;
                mbar_wait_b(bar_tmem_empty + 8 * acc, ((np >> 1) & 1) ^ 1);      // both CTAs drained this accumulator
                tc_fence_after();
                const uint32_t dcol = tmem_base + acc * 256;
                // Board rows -1 and 8 are zero padding: for the first / last row pair, three taps of one tile
                // multiply an all-zero operand and are not issued (both CTAs of the pair work on the same rows).
                const int yp = (pp / ntiles) % kRowPasses;
                const bool skip_top = yp == 0, skip_bot = yp == kRowPasses - 1;
                for (int kb = 0; kb < p.kblocks; ++kb, ++ait) {
                    const uint32_t as = ait % kActStagesP;
                    mbar_wait_b(bar_act_full + 8 * as, (ait / kActStagesP) & 1);
                    OZ_TL(3, lane == 0 && np == 0 && kb == 0);
                    // A K block's nine tap tiles are one ring group: ONE wait, one uninterrupted burst of MMAs, one commit.  (With
                    // a wait and a commit per tap the issue thread set the pace of the streamed-weight towers: 417 clocks per tap
                    // measured in the cluster tower against 256 clocks of tensor work, tensor pipe 66 % active at 256 channels.)
                    const uint32_t g = resident ? (uint32_t)kb : wgit % kWGroups;
                    if (!resident || np == 0) mbar_wait_b(bar_w_full + 8 * g, resident ? 0u : (wgit / kWGroups) & 1);
                    tc_fence_after();
                    const uint32_t abox = act0 + kSlotBytes + as * kActBoxBytesP;
                    if (elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                            const uint32_t wbase = w0 + (g * 9 + (uint32_t)tap) * kWTileBytes;
#pragma unroll
                            for (int r = 0; r < kYPerPass; ++r) {
                                if ((r == 0 && tap < 3 && skip_top) || (r == kYPerPass - 1 && tap >= 6 && skip_bot)) continue;
                                const int first_tap = (r == 0 && skip_top) ? 3 : 0;
                                const uint32_t arow = abox + (uint32_t)(((r + 1 + dy) * kBoxX + dx) * kSlotBytes);
#pragma unroll
                                for (int j = 0; j < kKBlock / 16; ++j)
                                    tc_mma_pair(dcol + r * 128, umma_desc_sw64(arow + j * 32), umma_desc_sw64(wbase + j * 32),
                                                kIdescP, (kb | j) != 0 || tap != first_tap);
                            }
                        }
                        tc_commit_pair(bar_act_empty + 8 * as);
                        if (!resident) tc_commit_pair(bar_w_empty + 8 * g);
                    }
                    __syncwarp();
                    ++wgit;
                }
                if (elect_one()) tc_commit_pair(bar_tmem_full + 8 * acc);
                __syncwarp();
                OZ_TL(4, lane == 0);
            }
        } else if (dynamic) {
            // ===== pass scheduler (the non-leader CTA's otherwise idle warp) =====
            for (uint32_t si = 0;; ++si) {
                // Entry si is drawn once the producer has read entry si - 2: the pair holds at most two undrawn passes, and the
                // round trip (report, atomic, publication: about 2 us) stays hidden even for the 1-K-block first layer.
                if (si > 1) mbar_wait_b(bar_sched_empty + 8 * ((si - 2) % kSchedSlots), ((si - 2) / kSchedSlots) & 1);
                int pp = 0;
                if (lane == 0) pp = atomicAdd(p.work, 1);
                pp = __shfl_sync(0xffffffffu, pp, 0);
                const uint32_t slot = si % kSchedSlots;
                if (elect_one()) {
                    st_cluster_b32(mapa_rank(sched_ring + 8 * slot, 0), pp);
                    st_cluster_b32(mapa_rank(sched_ring + 8 * slot, 1), pp);
                    mbar_arrive_cluster(mapa_rank(bar_sched_full + 8 * slot, 0));
                    mbar_arrive_cluster(mapa_rank(bar_sched_full + 8 * slot, 1));
                }
                __syncwarp();
                if (pp >= num_pp) break;
            }
        }
    } else {
        // ===== epilogue warps 3..10: TMEM lane quarter (warp % 4) x column half =====
        const int quad = warp & 3, half = (warp - 3) >> 2;
        const int row = quad * 32 + lane, x = row >> 4, bl = row & 15;     // row in tile = x * 16 + board
        const uint32_t empty_remote = mapa_rank(bar_tmem_empty, 0);
        float st_sum[3][2], st_sq[3][2];                      // STATS: [cout tile][32-channel chunk] totals of channel (... + lane)
#pragma unroll
        for (int a = 0; a < 3; ++a) { st_sum[a][0] = st_sum[a][1] = 0.f; st_sq[a][0] = st_sq[a][1] = 0.f; }
        griddep_wait();
        for (uint32_t np = 0;; ++np) {
            int pp;
            next_pass(np, pp);
            if (pp >= num_pp) break;
            const int nt = pp % ntiles, up = pp / ntiles, unit = 2 * (up / kRowPasses) + (int)rank, yp = up % kRowPasses;
            const uint32_t acc = np & 1;
            const size_t board = (size_t)unit * kBoardsPerUnit + bl;
            const size_t off0 = ((board * 8 + yp * kYPerPass) * 8 + x) * (size_t)p.cout_total + nt * 128 + half * 64;
            const size_t ystride = (size_t)8 * p.cout_total;
            uint32_t sk[STATS ? 1 : kYPerPass][32];            // 64 channels of residual input per tile row
            if (!STATS && p.skip) {
#pragma unroll
                for (int r = 0; r < kYPerPass; ++r) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) ldg256(p.skip + off0 + r * ystride + i * 16, &sk[r][i * 8]);
                }
                const int ppn = peek_pass(np, pp);
                if (ppn < num_pp) {                             // next pass's residual rows: HBM -> L2 ahead of use
                    const int ntn = ppn % ntiles, un = ppn / ntiles;
                    const size_t board_n = (size_t)(2 * (un / kRowPasses) + (int)rank) * kBoardsPerUnit + bl;
                    const size_t offn = ((board_n * 8 + (un % kRowPasses) * kYPerPass) * 8 + x) * (size_t)p.cout_total + ntn * 128 + half * 64;
#pragma unroll
                    for (int r = 0; r < kYPerPass; ++r)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.skip + offn + r * ystride));
                }
            }
            mbar_wait_b(bar_tmem_full + 8 * acc, (np >> 1) & 1);
            OZ_TL(5, threadIdx.x == 96);
            tc_fence_after();
#pragma unroll
            for (int r = 0; r < kYPerPass; ++r) {
                float h0 = 0.f, h1 = 0.f, h2 = 0.f;          // fused 1x1 head convolutions over this thread's 64 channels
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256 + r * 128 + half * 64 + ch * 32, v);
                    tc_wait_ld();
                    const float *bs = s_bias + nt * 128 + half * 64 + ch * 32;
                    uint32_t ow[16];
                    const uint32_t *skw = &sk[STATS ? 0 : r][ch * 16];
                    float fx[STATS ? 32 : 1], fq[STATS ? 32 : 1];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float a = __uint_as_float(v[2 * i]) + bs[2 * i];
                        float b = __uint_as_float(v[2 * i + 1]) + bs[2 * i + 1];
                        if (STATS) { fx[2 * i] = a; fx[2 * i + 1] = b; fq[2 * i] = a * a; fq[2 * i + 1] = b * b; }
                        if (!STATS && p.skip) {
                            const __nv_bfloat162 s2 = *(const __nv_bfloat162 *)&skw[i];
                            a += __bfloat162float(s2.x);
                            b += __bfloat162float(s2.y);
                        }
                        if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                        if (p.head_out) {
                            const int c = half * 64 + ch * 32 + 2 * i;
                            const float2 w0 = *(const float2 *)(s_head + c), w1 = *(const float2 *)(s_head + 128 + c), w2 = *(const float2 *)(s_head + 256 + c);
                            h0 += a * w0.x + b * w0.y; h1 += a * w1.x + b * w1.y; h2 += a * w2.x + b * w2.y;
                        }
                        const __nv_bfloat162 o2 = __floats2bfloat162_rn(a, b);
                        ow[i] = *(const uint32_t *)&o2;
                    }
                    if (p.out) {
                        __nv_bfloat16 *op = p.out + off0 + r * ystride + ch * 32;   // full 32-byte sectors per store
                        stg256(op, &ow[0]);
                        stg256(op + 16, &ow[8]);
                    }
                    if (STATS) {                             // 32 rows (lanes) x 32 channels -> lane j: channel j's sums
#pragma unroll
                        for (int sft = 16; sft >= 1; sft >>= 1) {
                            const bool upper = (lane & sft) != 0;
#pragma unroll
                            for (int i = 0; i < sft; ++i) {
                                const float sx = upper ? fx[i] : fx[i + sft], kx = upper ? fx[i + sft] : fx[i];
                                const float sq = upper ? fq[i] : fq[i + sft], kq = upper ? fq[i + sft] : fq[i];
                                fx[i] = kx + __shfl_xor_sync(0xffffffffu, sx, sft);
                                fq[i] = kq + __shfl_xor_sync(0xffffffffu, sq, sft);
                            }
                        }
#pragma unroll
                        for (int a = 0; a < 3; ++a)
                            if (a == nt) { st_sum[a][ch] += fx[0]; st_sq[a][ch] += fq[0]; }
                    }
                }
                if (p.head_out) {                            // the two column halves of a row add into zeroed memory: order-free
                    float *hp = p.head_out + ((board * 8 + yp * kYPerPass + r) * 8 + x) * 4;
                    atomicAdd(hp, h0); atomicAdd(hp + 1, h1); atomicAdd(hp + 2, h2);
                }
            }
            tc_fence_before();
            __syncwarp();
            OZ_TL(6, threadIdx.x == 96);
            if (lane == 0) mbar_arrive_cluster(empty_remote + 8 * acc);   // accumulator stage back to the leader's MMA thread
        }
        if (STATS && p.stats) {                               // this warp's 4 lane quarters x 32 rows each, all passes of the CTA
            float *dst = p.stats + (size_t)(blockIdx.x % p.stats_copies) * 2 * p.cout_total;
#pragma unroll
            for (int a = 0; a < 3; ++a)
                if (a < ntiles) {
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const int c = a * 128 + half * 64 + ch * 32 + lane;
                        atomicAdd(dst + c, st_sum[a][ch]);
                        atomicAdd(dst + p.cout_total + c, st_sq[a][ch]);
                    }
                }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
    OZ_TL(7, threadIdx.x == 32);
}

// ------------------------------------------------------------------ cluster tower (launch-bound batches)
// tower_cluster_kernel: EVERY convolution layer of a 128-channel forward over 32 boards in one launch of one 16-CTA
// cluster -- the batch the reference's own callers produce (`MCTS.search` of one game, T x B = 32 leaves per network
// call: train.py:415-416, player.py:251).  The per-layer path pays a kernel boundary per layer (about 10 us, of which
// 2.4 us are MMAs); a software grid barrier through global memory is no cheaper (profiles/r02_tower_experiment.md).
// Here the activations never leave the cluster's shared memory:
//   * CTA (y, u) = cluster rank 2 y + u owns board row y of unit u (16 boards): its 128-row M tile is the pass of
//     conv3x3_pair_kernel<1>, the pair (y, 0) / (y, 1) issues the same cta_group::2 MMAs in the same order (so the
//     result equals the per-layer path bit for bit).
//   * One activation buffer per CTA, 4 K blocks x (3 board rows x 9 x-slots) KB in the layout TMA would have written
//     (SWIZZLE_64B, row = (y * 9 + x) * 16 + board): row 1 is the CTA's own board row, rows 0 / 2 the halo.  The epilogue
//     writes a layer's bf16 output straight into row 1 of that layout; one warp then pushes the row, K block by K block,
//     into the halo rows of the CTAs above and below with bulk shared-to-shared copies (cp.async.bulk, 8 KB each, async
//     proxy end to end: no generic remote stores, no release fences).  The zero column x = 8, the guard slot and the rows
//     outside the board are zeroed once and never written.
//   * Layer hand-over is per K block and by mbarriers only, no grid or cluster barrier: `rows[kb]` counts the epilogue
//     warps that wrote K block kb of the own row, `halo[kb]` the bytes the neighbours' copies deliver, `peer[kb]` (leader)
//     one arrival from the other CTA of the pair once ITS rows[kb] and halo[kb] completed; the next layer's MMAs of K block
//     kb start as soon as those three completed, while the later K blocks are still in flight.  `nbr_done` receives the
//     tcgen05.commit of the neighbouring pairs' MMA threads (multicast to four CTAs): a row is only overwritten after the
//     neighbours' MMAs that read its copies.  The accumulator is double-buffered in TMEM (layer parity).
//   * The residual input of a block is the value the same epilogue thread produced two layers earlier: it stays in
//     32 registers (bf16 pairs), no memory at all.
//   * Weights: each CTA streams its 64-cout half of every (layer, K block) group of 9 tap tiles (36 KB) through a ring of
//     three groups, running ahead into the next layer; a K block is ONE barrier wait, one burst of 18 MMAs and one commit
//     (with a wait and a commit per tap the issue thread, not the tensor pipe, set the pace: 417 clocks per tap measured).
//   * The last layer's epilogue applies the 1x1 head convolutions (as the per-layer path does) -- no activation is ever
//     written to global memory.
// Several clusters = several independent 32-board groups (blockIdx.x / 16).
namespace tower {
constexpr int kCluster = 16;
constexpr int kBoxSlots = 3 * kBoxX;                                 // 27 slots: 3 board rows x 9
constexpr int kBoxBytes = kBoxSlots * kSlotBytes;                    // 27 648 per K block
constexpr int kRowCopyBytes = 8 * kSlotBytes;                        // x = 0..7 of one board row of one K block
constexpr int kActBytes = kSlotBytes + 4 * kBoxBytes;                // guard + 4 K blocks = 111 616
constexpr int kGroups = 3;                                           // weight ring: groups of 9 tap tiles
constexpr int kGroupBytes = 9 * pair::kWTileBytes;                   // 36 864
constexpr int kBarBytes = 512;
constexpr int kHeadBytes = 3 * 128 * 4;
constexpr int kBiasBytes = pair::kEpiWarps * 64 * 4;                 // one private 64-float slice per epilogue warp
constexpr int kSmemBytes = 1024 + kActBytes + kGroups * kGroupBytes + kBarBytes + kHeadBytes + kBiasBytes;
constexpr int kThreads = pair::kThreadsP;
static_assert(kSmemBytes <= 232448, "tower shared memory");
}  // namespace tower

struct TowerParams {
    const float *bias0;          // layer 0 bias [128]
    const float *bias1;          // layer 1 bias; layer l >= 1 at bias1 + (l - 1) * bias_stride floats
    size_t bias_stride;
    const float *head_w;         // [3][128]
    float *head_out;             // [boards][64][4] (every element of the cluster's 32 boards is stored: no zeroing needed)
    int nlayers;                 // 1 + 2 * blocks
    long long *dbg;              // timeline of cluster rank 8 (OZ_TOWER_DBG set): [layer][8] clock64 stamps, or null
};

__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_mask(uint32_t bar, uint16_t mask) {   // arrives on the same barrier offset in every CTA of `mask`
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint32_t *v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// Shared memory of this CTA -> shared memory of another CTA of the cluster; the bytes complete on a barrier of the destination CTA.
__device__ __forceinline__ void bulk_copy_to_cta(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(tower::kThreads, 1)
tower_cluster_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_w0,
                     const __grid_constant__ CUtensorMap map_wr, const TowerParams p) {
    using namespace tower;
    constexpr int kWTileBytes = pair::kWTileBytes;
    constexpr uint32_t kIdescP = pair::kIdescP;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t act0 = base;                                            // guard slot, then the 4 K-block boxes
    const uint32_t box1 = act0 + kSlotBytes;
    const uint32_t w0 = act0 + kActBytes;
    const uint32_t misc = w0 + kGroups * kGroupBytes;
    const uint32_t bar_in_full = misc, bar_tmem_full = misc + 8, bar_nbr_done = misc + 16;
    const uint32_t bar_rows = misc + 32, bar_halo = misc + 64, bar_peer = misc + 96;      // [4] each, per K block
    const uint32_t bar_w_full = misc + 128, bar_w_empty = misc + 160;                     // [kGroups] each
    uint32_t *tmem_ptr = (uint32_t *)(gen + (misc - base) + 192);
    float *s_head = (float *)(gen + (misc - base) + kBarBytes);
    float *s_bias = s_head + 3 * 128;                                      // [8 epilogue warps][64]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t lead_rank = rank & ~1u;
    const bool leader = rank == lead_rank;
    const int y = (int)(rank >> 1), u = (int)(rank & 1);
    const int unit = 2 * (int)(blockIdx.x / kCluster) + u;
    const int n_up = y > 0 ? 1 : 0, n_dn = y < 7 ? 1 : 0;
    const int L = p.nlayers;

    if (threadIdx.x == 0) {
        mbar_init(bar_in_full, 1);
        mbar_init(bar_tmem_full, 1);
        mbar_init(bar_nbr_done, n_up + n_dn);
        for (int i = 0; i < 4; ++i) { mbar_init(bar_rows + 8 * i, 4); mbar_init(bar_halo + 8 * i, 1); mbar_init(bar_peer + 8 * i, 1); }
        for (int i = 0; i < kGroups; ++i) { mbar_init(bar_w_full + 8 * i, 1); mbar_init(bar_w_empty + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wr) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kActBytes / 16; i += kThreads)           // zeros: guard, x = 8 column, rows outside the board
        *(uint4 *)(gen + i * 16) = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 3 * 128; i += kThreads) s_head[i] = p.head_w[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    griddep_launch_dependents();

    if (warp == 0) {
        // ===== weight TMA producer: this CTA's 64-cout half of every (layer, K block) group of 9 tap tiles, in MMA order =====
        const uint32_t full0 = mapa_rank(bar_w_full, lead_rank);
        uint32_t git = 0;
        for (int l = 0; l < L; ++l) {
            const int kblocks = l == 0 ? 1 : 4;
            for (int kb = 0; kb < kblocks; ++kb, ++git) {
                const uint32_t g = git % kGroups, ph = (git / kGroups) & 1;
                mbar_wait_b(bar_w_empty + 8 * g, ph ^ 1);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(bar_w_full + 8 * g, 2 * kGroupBytes);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t dst = w0 + g * kGroupBytes + tap * kWTileBytes;
                        if (l == 0) tma_load_2d_pair(dst, &map_w0, full0 + 8 * g, 0, tap * 128 + u * 64);
                        else tma_load_3d_pair(dst, &map_wr, full0 + 8 * g, kb * kKBlock, tap * 128 + u * 64, l - 1);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ===== network input, then the halo pushes: own board row -> row 2 of the CTA above, row 0 of the CTA below =====
        griddep_wait();
        if (elect_one()) {
            if (leader) mbar_expect_tx(bar_in_full, 2 * kBoxBytes);
            tma_load_4d_pair(box1, &map_x0, mapa_rank(bar_in_full, lead_rank), 0, unit * kBoardsPerUnit, 0, y - 1);
        }
        __syncwarp();
        const uint32_t src_row = box1 + kBoxX * kSlotBytes;
        const uint32_t up_row = n_up ? mapa_rank(box1 + 2 * kBoxX * kSlotBytes, rank - 2) : 0u;
        const uint32_t dn_row = n_dn ? mapa_rank(box1, rank + 2) : 0u;
        const uint32_t up_bar = n_up ? mapa_rank(bar_halo, rank - 2) : 0u, dn_bar = n_dn ? mapa_rank(bar_halo, rank + 2) : 0u;
        for (int l = 0; l + 1 < L; ++l) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int kb = ((k & 1) << 1) | (k >> 1);                    // the order the epilogue finishes them: 0, 2, 1, 3
                mbar_wait_b(bar_rows + 8 * kb, (uint32_t)l & 1);             // written, proxy-fenced; the neighbours' MMAs of layer l are done
                if (elect_one()) {
                    mbar_expect_tx(bar_halo + 8 * kb, (uint32_t)(n_up + n_dn) * kRowCopyBytes);   // what the neighbours send us
                    const uint32_t off = (uint32_t)kb * kBoxBytes;
                    if (n_up) bulk_copy_to_cta(up_row + off, src_row + off, kRowCopyBytes, up_bar + 8 * kb);
                    if (n_dn) bulk_copy_to_cta(dn_row + off, src_row + off, kRowCopyBytes, dn_bar + 8 * kb);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===== MMA issuer (pair leader) =====
            const uint16_t pair_mask = (uint16_t)(3u << lead_rank);
            const uint16_t nbr_mask = (uint16_t)((n_up ? (3u << (lead_rank - 2)) : 0u) | (n_dn ? (3u << (lead_rank + 2)) : 0u));
            const bool skip_top = y == 0, skip_bot = y == 7;                 // rows -1 / 8 are zero padding: taps not issued
            const int first_tap = skip_top ? 3 : 0;
            uint32_t git = 0;
            for (int l = 0; l < L; ++l) {
                const uint32_t dcol = tmem_base + (uint32_t)(l & 1) * 128;
                const int kblocks = l == 0 ? 1 : 4;
                for (int kb = 0; kb < kblocks; ++kb, ++git) {
                    if (l == 0) mbar_wait_b(bar_in_full, 0);
                    else {
                        const uint32_t ph = (uint32_t)(l - 1) & 1;
                        mbar_wait_b(bar_rows + 8 * kb, ph);                  // own row, K block kb
                        if (p.dbg && rank == 8 && lane == 0 && kb == 0) p.dbg[l * 8 + 6] = clock64();
                        mbar_wait_b(bar_halo + 8 * kb, ph);                  // halo rows from above / below
                        if (p.dbg && rank == 8 && lane == 0 && kb == 0) p.dbg[l * 8 + 7] = clock64();
                        mbar_wait_cluster(bar_peer + 8 * kb, ph);            // the same two in the other CTA of the pair
                    }
                    const uint32_t g = git % kGroups;
                    mbar_wait_b(bar_w_full + 8 * g, (git / kGroups) & 1);
                    tc_fence_after();
                    if (p.dbg && rank == 8 && lane == 0 && kb == 0) p.dbg[l * 8 + 0] = clock64();
                    const uint32_t abox = box1 + (uint32_t)kb * kBoxBytes;
                    if (elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            if ((tap < 3 && skip_top) || (tap >= 6 && skip_bot)) continue;
                            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                            const uint32_t wbase = w0 + g * kGroupBytes + tap * kWTileBytes;
                            const uint32_t arow = abox + (uint32_t)(((1 + dy) * kBoxX + dx) * kSlotBytes);
#pragma unroll
                            for (int j = 0; j < kKBlock / 16; ++j)
                                tc_mma_pair(dcol, umma_desc_sw64(arow + j * 32), umma_desc_sw64(wbase + j * 32),
                                            kIdescP, (kb | j) != 0 || tap != first_tap);
                        }
                        tc_commit_mask(bar_w_empty + 8 * g, pair_mask);
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    tc_commit_mask(bar_tmem_full, pair_mask);
                    if (l + 1 < L && nbr_mask) tc_commit_mask(bar_nbr_done, nbr_mask);   // the neighbours may overwrite our halo rows
                }
                __syncwarp();
                if (p.dbg && rank == 8 && lane == 0) p.dbg[l * 8 + 1] = clock64();
            }
        } else {
            // ===== relay (the pair's other CTA): tell the leader when K block kb of OUR buffer is complete =====
            const uint32_t peer0 = mapa_rank(bar_peer, lead_rank);
            for (int l = 0; l + 1 < L; ++l)
                for (int kb = 0; kb < 4; ++kb) {
                    mbar_wait_b(bar_rows + 8 * kb, (uint32_t)l & 1);
                    mbar_wait_b(bar_halo + 8 * kb, (uint32_t)l & 1);
                    // relaxed: what the arrival announces is already in this CTA's shared memory (the bulk copies' bytes completed on
                    // `halo`, the epilogue's stores were proxy-fenced before `rows`); a release here costs about 600 clocks
                    if (lane == 0) asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(peer0 + 8 * kb) : "memory");
                    __syncwarp();
                }
        }
    } else if (warp >= 3) {
        // ===== epilogue warps 3..10: TMEM lane quarter (warp % 4) x column half =====
        const int quad = warp & 3, half = (warp - 3) >> 2;
        const int row = quad * 32 + lane, x = row >> 4, bl = row & 15;        // row in tile = x * 16 + board
        const uint32_t sw = (uint32_t)(bl >> 1) & 3u;                         // SWIZZLE_64B: 16-byte chunk ^= address bits 7..8
        const uint32_t row_addr = box1 + (uint32_t)((kBoxX + x) * kSlotBytes + bl * kRowBytes);   // own board row = row 1 of a box
        const size_t board = (size_t)unit * kBoardsPerUnit + bl;
        float *bias_w = s_bias + (warp - 3) * 64;                             // this warp's 64 bias values of the current layer
        uint32_t keep[32];                                                    // block input of this thread's 64 channels (bf16 pairs)
#pragma unroll
        for (int i = 0; i < 32; ++i) keep[i] = 0u;
        griddep_wait();
        for (int l = 0; l < L; ++l) {
            const bool last = l + 1 == L;
            const bool residual = l >= 2 && (l & 1) == 0;                     // second convolution of a block (neural_net.py:57-64)
            if (lane < 16) {
                const float *bsrc = (l == 0 ? p.bias0 : p.bias1 + (size_t)(l - 1) * p.bias_stride) + half * 64 + lane * 4;
                *(float4 *)(bias_w + lane * 4) = __ldg((const float4 *)bsrc);
            }
            __syncwarp();
            mbar_wait_b(bar_tmem_full, (uint32_t)l & 1);
            const bool stamp = p.dbg && rank == 8 && warp == 3 && lane == 0;
            if (stamp) p.dbg[l * 8 + 2] = clock64();
            tc_fence_after();
            float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(l & 1) * 128 + half * 64 + ch * 32, v);
                tc_wait_ld();
                const float *bs = bias_w + ch * 32;
                uint32_t ow[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float a = __uint_as_float(v[2 * i]) + bs[2 * i];
                    float b = __uint_as_float(v[2 * i + 1]) + bs[2 * i + 1];
                    if (residual) {
                        const __nv_bfloat162 s2 = *(const __nv_bfloat162 *)&keep[ch * 16 + i];
                        a += __bfloat162float(s2.x);
                        b += __bfloat162float(s2.y);
                    }
                    a = fmaxf(a, 0.f); b = fmaxf(b, 0.f);
                    if (last) {
                        const int c = half * 64 + ch * 32 + 2 * i;
                        const float2 w0h = *(const float2 *)(s_head + c), w1h = *(const float2 *)(s_head + 128 + c), w2h = *(const float2 *)(s_head + 256 + c);
                        h0 += a * w0h.x + b * w0h.y; h1 += a * w1h.x + b * w1h.y; h2 += a * w2h.x + b * w2h.y;
                    }
                    const __nv_bfloat162 o2 = __floats2bfloat162_rn(a, b);
                    ow[i] = *(const uint32_t *)&o2;
                    if ((l & 1) == 0) keep[ch * 16 + i] = ow[i];             // a block's output is the next block's residual input
                }
                if (!last) {
                    // The row we overwrite is the source of the previous layer's pushes and what the neighbours' MMAs of THIS
                    // layer read as halo (through their copies): both are over once the neighbours' layer-l MMAs committed.
                    if (ch == 0) mbar_wait_b(bar_nbr_done, (uint32_t)l & 1);
                    if (stamp && ch == 0) p.dbg[l * 8 + 3] = clock64();
                    const int kb = half * 2 + ch;
                    const uint32_t dst = row_addr + (uint32_t)kb * kBoxBytes;
#pragma unroll
                    for (int q = 0; q < 4; ++q) st_shared_v4(dst + (((uint32_t)q ^ sw) * 16u), &ow[4 * q]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> bulk copy / tensor-core reads
                    if (ch == 1) tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_local(bar_rows + 8 * kb);
                    if (stamp) p.dbg[l * 8 + 4 + ch] = clock64();
                }
            }
            if (last) {
                // The per-layer path adds the two column halves of a row into zeroed memory with atomics (a + b in either order:
                // the same bits).  Here the upper half hands its sums over through shared memory (the activation buffer is free
                // after the last MMA) and the lower half stores the total: no atomics, and head_out needs no zeroing.
                float4 *xch = (float4 *)(gen + kSlotBytes) + row;
                if (half == 1) *xch = make_float4(h0, h1, h2, 0.f);
                asm volatile("bar.sync 1, 256;" ::: "memory");                // the eight epilogue warps
                if (half == 0) {
                    const float4 o = *xch;
                    *(float4 *)(p.head_out + ((board * 8 + y) * 8 + x) * 4) = make_float4(h0 + o.x, h1 + o.y, h2 + o.z, 0.f);
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------ input conversion
// float planes [n][planes][64] -> bf16 NHWC [n_pad][64][32] (channels >= planes are zero).
__global__ void planes_to_nhwc32_kernel(const float *__restrict__ f, int planes, size_t n, size_t n_pad, __nv_bfloat16 *__restrict__ out) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;   // one thread per (board, pos, 8-channel group)
    if (idx >= n_pad * 64 * 4) return;
    const size_t b = idx / 256;
    const int pos = (int)((idx / 4) % 64), g = (int)(idx % 4);
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = g * 8 + i;
        v[i] = __float2bfloat16((b < n && c < planes) ? f[(b * planes + c) * 64 + pos] : 0.f);
    }
    *(uint4 *)(out + (b * 64 + pos) * 32 + g * 8) = *(const uint4 *)v;
}

// ------------------------------------------------------------------ heads
struct HeadParams {
    const __nv_bfloat16 *act;  // [n][64][C]
    const float *pconv_w, *pconv_b, *pfc_w, *pfc_b, *vconv_w, *vconv_b, *vfc1_w, *vfc1_b, *vfc2_w, *vfc2_b;
    const float *pre;          // optional [n][64][4]: 1x1 head-conv pre-activations written by the last conv layer
    float *policy, *value;
    int C, Hv;
    size_t n;
};

// neural_net.py:68-128.  One WARP works on kHeadBoards boards at a time (8 warps per CTA): lane l owns board squares l
// and l + 32 for the two 1x1 convolutions, then policy outputs l and l + 32 (output 64 is a warp reduction) and every
// 32nd hidden unit of the value head.  FC weights live in shared memory (row strides = 4 mod 32 floats: conflict-free
// 128-bit loads) and every weight vector read is used for all kHeadBoards boards, so the kernel is no longer bound by
// shared-memory bandwidth (the one-board-per-warp version did 3 LDS per 2 FMA).  Summation order per output is the plain
// i = 0, 1, 2, ... order.
// Small batches (the single-game search: 32 boards; 256x20b at leaf 64 ... 1 024) run the <1> instantiation: one board
// per warp, four times as many warps in flight, a quarter of the serial work per warp -- the kernel is then a latency
// chain, not a throughput problem.  Every output is the same sum in the same order in both instantiations (the 1x1
// convolutions are written as explicit fused multiply-adds for that reason), so which one runs is invisible in the bits.
constexpr int kHeadWarps = 8;
constexpr int kHeadBoardsWide = 4;
constexpr int kPfcStride = 132, kVfcStride = 68, kFeatStride = 192;   // floats
__host__ __device__ inline int heads_cw_offset(int Hv) { return (65 * kPfcStride + Hv * kVfcStride + 3) & ~3; }       // floats
__host__ __device__ inline int heads_smem_floats(int C, int Hv, int boards_per_warp) {
    return ((heads_cw_offset(Hv) + 3 * C + 2 * Hv + 3) & ~3) + kHeadWarps * boards_per_warp * kFeatStride;
}

template <int kHeadBoards>
__global__ void __launch_bounds__(kHeadWarps * 32, 2) heads_kernel(const HeadParams p) {
    extern __shared__ __align__(16) float sm[];
    const int C = p.C, Hv = p.Hv;
    float *s_pfc = sm;                               // [65][132]
    float *s_vfc1 = s_pfc + 65 * kPfcStride;         // [Hv][68]
    float *s_cw = sm + heads_cw_offset(Hv);          // [3][C]: policy c0, policy c1, value (16-byte aligned)
    float *s_vb1 = s_cw + 3 * C;                     // [Hv] value fc1 bias
    float *s_vw2 = s_vb1 + Hv;                       // [Hv] value fc2 weight
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // per warp and board: [128] policy features (channel-major), then [64] value features
    float *s_feat = sm + ((heads_cw_offset(Hv) + 3 * C + 2 * Hv + 3) & ~3) + warp * kHeadBoards * kFeatStride;
    // Behind the cluster tower this kernel is launched with programmatic dependent launch: the weight staging below (constant
    // data) runs while the tower still computes; the activations are touched after griddep_wait().
    griddep_launch_dependents();
    {   // 128-bit loads; both row strides are multiples of 4 floats.  All of a thread's loads are issued before the first
        // store: one L2 round trip instead of nine (this loop was not unrolled and cost 6 us wherever the kernel cannot
        // start under its predecessor -- every per-layer forward, whose last convolution fills the SMs' shared memory)
        constexpr int kPfcIters = (65 * 32 + kHeadWarps * 32 - 1) / (kHeadWarps * 32);
        float4 stage[kPfcIters];
#pragma unroll
        for (int k = 0; k < kPfcIters; ++k) {
            const int i = t + k * kHeadWarps * 32;
            if (i < 65 * 32) stage[k] = __ldg((const float4 *)p.pfc_w + i);
        }
#pragma unroll
        for (int k = 0; k < kPfcIters; ++k) {
            const int i = t + k * kHeadWarps * 32;
            if (i < 65 * 32) *(float4 *)(s_pfc + (i >> 5) * kPfcStride + (i & 31) * 4) = stage[k];
        }
    }
    for (int i = t; i < Hv * 16; i += kHeadWarps * 32)
        *(float4 *)(s_vfc1 + (i >> 4) * kVfcStride + (i & 15) * 4) = __ldg((const float4 *)p.vfc1_w + i);
    for (int i = t; i < 2 * C; i += kHeadWarps * 32) s_cw[i] = p.pconv_w[i];
    for (int i = t; i < C; i += kHeadWarps * 32) s_cw[2 * C + i] = p.vconv_w[i];
    for (int i = t; i < Hv; i += kHeadWarps * 32) { s_vb1[i] = p.vfc1_b[i]; s_vw2[i] = p.vfc2_w[i]; }
    __syncthreads();
    griddep_wait();
    const float pb0 = p.pconv_b[0], pb1 = p.pconv_b[1], vb = p.vconv_b[0], v2b = p.vfc2_b[0];
    const float pfb0 = p.pfc_b[lane], pfb1 = p.pfc_b[lane + 32], pfb64 = p.pfc_b[64];
    const float4 *cw0 = (const float4 *)s_cw, *cw1 = (const float4 *)(s_cw + C), *cw2 = (const float4 *)(s_cw + 2 * C);
    for (size_t b0 = ((size_t)blockIdx.x * kHeadWarps + warp) * kHeadBoards; b0 < p.n; b0 += (size_t)gridDim.x * kHeadWarps * kHeadBoards) {
        // ---- 1x1 convolutions + BN + ReLU (or their fused pre-activations) -> features in shared memory
#pragma unroll
        for (int bi = 0; bi < kHeadBoards; ++bi) {
            const size_t b = b0 + bi < p.n ? b0 + bi : p.n - 1;      // tail: recompute the last board, never stored
            float *feat = s_feat + bi * kFeatStride, *v1 = feat + 128;
            if (p.pre) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int sq = lane + 32 * h;
                    const float4 d = __ldg((const float4 *)(p.pre + (b * 64 + sq) * 4));
                    feat[sq] = fmaxf(d.x + pb0, 0.f);
                    feat[64 + sq] = fmaxf(d.y + pb1, 0.f);
                    v1[sq] = fmaxf(d.z + vb, 0.f);
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int sq = lane + 32 * h;
                    const __nv_bfloat16 *a = p.act + (b * 64 + sq) * (size_t)C;
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll (kHeadBoards == 1 ? 4 : 2)
                    for (int c16 = 0; c16 < C / 16; ++c16) {        // one full 32-byte sector (16 channels) per load
                        uint32_t q[8];
                        ldg256(a + c16 * 16, q);
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const __nv_bfloat162 *hh = (const __nv_bfloat162 *)&q[4 * h2];
                            const float x0 = __bfloat162float(hh[0].x), x1 = __bfloat162float(hh[0].y), x2 = __bfloat162float(hh[1].x), x3 = __bfloat162float(hh[1].y);
                            const float x4 = __bfloat162float(hh[2].x), x5 = __bfloat162float(hh[2].y), x6 = __bfloat162float(hh[3].x), x7 = __bfloat162float(hh[3].y);
                            const int c8 = 2 * c16 + h2;
                            // eight channels: one product, seven explicit fmas, then one add into the running sum
                            auto dot8 = [&](const float4 wa, const float4 wb) {
                                float s8 = __fmul_rn(x0, wa.x);
                                s8 = __fmaf_rn(x1, wa.y, s8); s8 = __fmaf_rn(x2, wa.z, s8); s8 = __fmaf_rn(x3, wa.w, s8);
                                s8 = __fmaf_rn(x4, wb.x, s8); s8 = __fmaf_rn(x5, wb.y, s8); s8 = __fmaf_rn(x6, wb.z, s8);
                                return __fmaf_rn(x7, wb.w, s8);
                            };
                            d0 = __fadd_rn(d0, dot8(cw0[2 * c8], cw0[2 * c8 + 1]));
                            d1 = __fadd_rn(d1, dot8(cw1[2 * c8], cw1[2 * c8 + 1]));
                            d2 = __fadd_rn(d2, dot8(cw2[2 * c8], cw2[2 * c8 + 1]));
                        }
                    }
                    feat[sq] = fmaxf(d0 + pb0, 0.f);
                    feat[64 + sq] = fmaxf(d1 + pb1, 0.f);
                    v1[sq] = fmaxf(d2 + vb, 0.f);
                }
            }
        }
        __syncwarp();
        // ---- policy FC 128 -> 65: outputs lane, lane + 32 per board; output 64 by warp reduction
        float l0[kHeadBoards], l1[kHeadBoards], l2[kHeadBoards];
#pragma unroll
        for (int bi = 0; bi < kHeadBoards; ++bi) { l0[bi] = pfb0; l1[bi] = pfb1; l2[bi] = 0.f; }
        {
            const float4 *w0 = (const float4 *)(s_pfc + lane * kPfcStride), *w1 = (const float4 *)(s_pfc + (lane + 32) * kPfcStride);
#pragma unroll 4
            for (int i4 = 0; i4 < 32; ++i4) {
                const float4 a = w0[i4], c = w1[i4];
#pragma unroll
                for (int bi = 0; bi < kHeadBoards; ++bi) {
                    const float4 f = *(const float4 *)(s_feat + bi * kFeatStride + 4 * i4);
                    l0[bi] += f.x * a.x; l0[bi] += f.y * a.y; l0[bi] += f.z * a.z; l0[bi] += f.w * a.w;
                    l1[bi] += f.x * c.x; l1[bi] += f.y * c.y; l1[bi] += f.z * c.z; l1[bi] += f.w * c.w;
                }
            }
            const float4 w2 = *(const float4 *)(s_pfc + 64 * kPfcStride + lane * 4);
#pragma unroll
            for (int bi = 0; bi < kHeadBoards; ++bi) {
                const float4 f = *(const float4 *)(s_feat + bi * kFeatStride + 4 * lane);
                float x = 0.f;
                x += f.x * w2.x; x += f.y * w2.y; x += f.z * w2.z; x += f.w * w2.w;
#pragma unroll
                for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                l2[bi] = x + pfb64;
            }
        }
        // ---- value FC 64 -> Hv -> 1: hidden units lane + 32 j, four j at a time
        float hv[kHeadBoards];
#pragma unroll
        for (int bi = 0; bi < kHeadBoards; ++bi) hv[bi] = 0.f;
        for (int k0 = 0; k0 < Hv; k0 += 128) {
            float acc[kHeadBoards][4];
            const float4 *wr[4];
            bool live[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + lane + 32 * j;
                live[j] = k < Hv;
                wr[j] = (const float4 *)(s_vfc1 + (live[j] ? k : 0) * kVfcStride);
#pragma unroll
                for (int bi = 0; bi < kHeadBoards; ++bi) acc[bi][j] = live[j] ? s_vb1[k] : 0.f;
            }
#pragma unroll 2
            for (int i4 = 0; i4 < 16; ++i4) {
                float4 f[kHeadBoards];
#pragma unroll
                for (int bi = 0; bi < kHeadBoards; ++bi) f[bi] = *(const float4 *)(s_feat + bi * kFeatStride + 128 + 4 * i4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w = wr[j][i4];
#pragma unroll
                    for (int bi = 0; bi < kHeadBoards; ++bi) {
                        acc[bi][j] += f[bi].x * w.x; acc[bi][j] += f[bi].y * w.y; acc[bi][j] += f[bi].z * w.z; acc[bi][j] += f[bi].w * w.w;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float w2v = live[j] ? s_vw2[k0 + lane + 32 * j] : 0.f;
#pragma unroll
                for (int bi = 0; bi < kHeadBoards; ++bi) hv[bi] += fmaxf(acc[bi][j], 0.f) * w2v;
            }
        }
        // ---- softmax over 65 logits, tanh
#pragma unroll
        for (int bi = 0; bi < kHeadBoards; ++bi) {
            float h = hv[bi];
#pragma unroll
            for (int o = 16; o; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
            float m = fmaxf(fmaxf(l0[bi], l1[bi]), l2[bi]);
#pragma unroll
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const float e0 = expf(l0[bi] - m), e1 = expf(l1[bi] - m), e2 = expf(l2[bi] - m);
            float ssum = e0 + e1;
#pragma unroll
            for (int o = 16; o; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
            const float inv = 1.f / (ssum + e2);
            const size_t b = b0 + bi;
            if (b < p.n) {
                float *po = p.policy + b * 65;
                po[lane] = e0 * inv; po[lane + 32] = e1 * inv;
                if (lane == 0) { po[64] = e2 * inv; p.value[b] = tanhf(h + v2b); }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ host side

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

ResNetLayout resnet_layout(const oz_resnet_shape &s) {
    ResNetLayout L;
    size_t o = 0;
    const size_t C = s.channels, Hv = s.value_hidden;
    const int nconv = 1 + 2 * s.blocks;
    for (int i = 0; i < nconv; ++i) {
        const size_t cin = (i == 0) ? 32 : C;
        L.conv_w.push_back(o); o = align256(o + 9 * C * cin * 2);
        L.conv_b.push_back(o); o = align256(o + C * 4);
    }
    L.pconv_w = o; o = align256(o + 2 * C * 4);
    L.pconv_b = o; o = align256(o + 2 * 4);
    L.pfc_w = o; o = align256(o + 65 * 128 * 4);
    L.pfc_b = o; o = align256(o + 65 * 4);
    L.vconv_w = o; o = align256(o + C * 4);
    L.vconv_b = o; o = align256(o + 4);
    L.vfc1_w = o; o = align256(o + Hv * 64 * 4);
    L.vfc1_b = o; o = align256(o + Hv * 4);
    L.vfc2_w = o; o = align256(o + Hv * 4);
    L.vfc2_b = o; o = align256(o + 4);
    L.head_w = o; o = align256(o + 3 * C * 4);     // device-side scratch: [policy c0 | policy c1 | value] rows, filled by load()
    L.total = o;
    return L;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode(EncodeTiledFn *fn) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        OZ_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (!ptr || qres != cudaDriverEntryPointSuccess) return fail(OZ_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        cached = (EncodeTiledFn)ptr;
    }
    *fn = cached;
    return OZ_OK;
}

// NHWC [boards][8][8][C] bf16 viewed as (channel, board, x, y); box = (32, 16, 9, 6).
int make_act_map(CUtensorMap *m, void *ptr, size_t boards, int C, int box_y, int box_x) {
    EncodeTiledFn enc;
    OZ_TRY(get_encode(&enc));
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)boards, 8, 8};
    cuuint64_t strides[3] = {(cuuint64_t)64 * C * 2, (cuuint64_t)C * 2, (cuuint64_t)8 * C * 2};
    cuuint32_t box[4] = {kKBlock, kBoardsPerUnit, (cuuint32_t)box_x, (cuuint32_t)box_y};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OZ_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed: " + std::to_string((int)r));
    return OZ_OK;
}

// [tap][cout][cin] bf16 viewed as (cin, tap*cout); box = (32, 128).
static int make_w_map(CUtensorMap *m, void *ptr, int cin, int cout, int box_rows) {
    EncodeTiledFn enc;
    OZ_TRY(get_encode(&enc));
    cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)9 * cout};
    cuuint64_t strides[1] = {(cuuint64_t)cin * 2};
    cuuint32_t box[2] = {kKBlock, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OZ_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
    return OZ_OK;
}

// Residual-layer weights of a 128-channel blob as ONE 3-D map (cin, tap * cout, layer): the layers lie at a constant pitch
// (weights + bias, both 256-byte aligned), so the cluster tower addresses a tile by (K block, tap row, layer).
static int make_w_map_layers(CUtensorMap *m, void *ptr, int C, int layers, size_t layer_pitch_bytes) {
    EncodeTiledFn enc;
    OZ_TRY(get_encode(&enc));
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)9 * C, (cuuint64_t)layers};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)layer_pitch_bytes};
    cuuint32_t box[3] = {kKBlock, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OZ_ERR_CUDA, "cuTensorMapEncodeTiled(weights, layers) failed: " + std::to_string((int)r));
    return OZ_OK;
}

// Clusters are independent 32-board groups, so a forward may hold more of them than the device can run at once: two waves of
// the tower (2 x 85 us at 128x10b) still beat the per-layer launches (230 us at 256 boards); beyond that they do not.
constexpr int kTowerWaves = 2;
// OZ_TOWER=0 in the environment keeps launch-bound batches on the per-layer path (A/B runs).
long long *g_tower_dbg = nullptr;
#ifdef OZ_CONV_TIMELINE
unsigned long long *g_conv_tl = nullptr;
extern "C" int oz_debug_conv_timeline(unsigned long long *out) { cudaDeviceSynchronize(); return g_conv_tl ? (int)cudaMemcpy(out, g_conv_tl, 128 * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) : -1; }
#endif
long long *tower_dbg_ptr() { return g_tower_dbg; }
static int g_tower_mode = -1;                                // -1: environment, 0 / 1: set by oz_debug_set_tower
void set_tower_mode(int mode) { g_tower_mode = mode < 0 ? -1 : (mode ? 1 : 0); }
static int g_heads_boards = 0;                               // 0: by batch size, 1 / 4: set by oz_debug_set_heads_boards
void set_heads_boards(int boards) { g_heads_boards = boards == 1 ? 1 : (boards == kHeadBoardsWide ? kHeadBoardsWide : 0); }
static int heads_boards_override() {
    if (g_heads_boards) return g_heads_boards;
    static int v = -1;
    if (v < 0) { const char *e = getenv("OZ_HEADS_BOARDS"); v = e ? atoi(e) : 0; if (v != 1 && v != kHeadBoardsWide) v = 0; }
    return v;
}
static bool tower_enabled() {
    if (g_tower_mode >= 0) return g_tower_mode == 1;
    static int v = -1;
    if (v < 0) { const char *e = getenv("OZ_TOWER"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

// How many 16-CTA clusters of tower_cluster_kernel the current device can hold at once (0: the tower is not used).
int tower_max_clusters() {
    static int cached[64] = {};                              // per device ordinal: -1 unusable, > 0 clusters
    int dev = 0;
    cudaGetDevice(&dev);
    int &c = cached[dev & 63];
    if (c != 0) return c > 0 ? c : 0;
    c = -1;
    const int smem = tower::kSmemBytes;
    if (cudaFuncSetAttribute(tower_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tower::kSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(tower_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tower::kCluster); cfg.blockDim = dim3(tower::kThreads); cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = tower::kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, tower_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (n > 0) c = n;
    return n;
}

bool ResNet::tower_capable() const { return has_wr_ && shape_.channels == 128 && shape_.blocks > 0 && tower_enabled() && tower_max_clusters() > 0; }

int ResNet::launch_tower(const Workspace &w, size_t pad, cudaStream_t stream) {
    const int nconv = 1 + 2 * shape_.blocks;
    TowerParams tp{};
    tp.bias0 = (const float *)(d_blob_ + layout_.conv_b[0]);
    tp.bias1 = (const float *)(d_blob_ + layout_.conv_b[1]);
    tp.bias_stride = nconv > 2 ? (layout_.conv_b[2] - layout_.conv_b[1]) / sizeof(float) : 0;
    tp.head_w = (const float *)(d_blob_ + layout_.head_w);
    tp.head_out = w.head_pre;
    tp.nlayers = nconv;
    {
        static long long *dbg = nullptr;
        const char *e = getenv("OZ_TOWER_DBG");
        if (e && !dbg) { cudaMalloc(&dbg, 64 * 8 * sizeof(long long)); cudaMemset(dbg, 0, 64 * 8 * sizeof(long long)); }
        tp.dbg = dbg;
        g_tower_dbg = dbg;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(pad / 32) * tower::kCluster); cfg.blockDim = dim3(tower::kThreads);
    cfg.dynamicSmemBytes = (size_t)tower::kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = tower::kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_ ? 2 : 1;
    OZ_CUDA(cudaLaunchKernelEx(&cfg, tower_cluster_kernel, w.map_x0_1, map_w_[0], map_wr_, tp));
    count_launch();
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

// ma = activation map with a 4-row box (two board rows per pass), ma1 = the same tensor with a 3-row box (one row per pass)
// or null.  One row per pass is used when that still leaves every CTA pair with at most one pass: the pass is half as
// long and twice as many pairs work.  (Beyond that it only costs: a pair would run two half passes instead of one, and
// with streamed weights -- 256 channels -- load every weight tile twice.)
static int launch_conv(const CUtensorMap &ma, const CUtensorMap *ma1, const CUtensorMap &mw, const ConvParams &cp, cudaStream_t stream,
                       bool pdl = true) {
    static unsigned long long attr_devices = 0;
    if (first_use_on_device(attr_devices)) {
        OZ_CUDA(cudaFuncSetAttribute(conv3x3_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::kSmemBytesP));
        OZ_CUDA(cudaFuncSetAttribute(conv3x3_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::kSmemBytesP));
        OZ_CUDA(cudaFuncSetAttribute(conv3x3_pair_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::kSmemBytesP));
        OZ_CUDA(cudaFuncSetAttribute(conv3x3_pair_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::kSmemBytesP));
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int pp2 = cp.units * 2 * (cp.cout_total >> 7);
    const bool one_row = ma1 != nullptr && 2 * pp2 <= sms / 2;   // every pair still gets at most one pass, half as long
    const int num_pp = one_row ? 2 * pp2 : pp2;
    // OZ_CONV_CLUSTERS (experiment switch): cap the persistent grid, e.g. 73 leaves one TPC free for the wide tree kernels of
    // the pipelined search; with the dynamic pass scheduler the convolution is indifferent to its grid size.
    static int cap = -1;
    if (cap < 0) { const char *e = getenv("OZ_CONV_CLUSTERS"); cap = e ? atoi(e) : 0; }
    const int max_clusters = (cap > 0 && cap < sms / 2) ? cap : sms / 2;
    const int clusters = num_pp < max_clusters ? num_pp : max_clusters;
    // The dynamic pass scheduler pays off when clusters run many passes (balance; robustness against co-running kernels); its
    // first index costs a counter round trip plus a cluster-wide publication (about 1.5 us) before any role can start, which
    // a launch with a pass or two per cluster -- the launch-bound small leaf batches -- never earns back: static there.
    ConvParams cpl = cp;
    if (num_pp < 8 * clusters) cpl.work = nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(pair::kThreadsP); cfg.dynamicSmemBytes = pair::kSmemBytesP; cfg.stream = stream;
    cudaLaunchAttribute attr[1];                       // programmatic dependent launch: see griddep_wait() in the kernel
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    if (cpl.stats) {
        if (cpl.skip || cpl.cout_total > 384) return fail(OZ_ERR_INVALID_ARGUMENT, "conv3x3 with statistics: no residual input, at most 384 channels");
        if (one_row) OZ_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel<1, true>, *ma1, mw, cpl));
        else OZ_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel<2, true>, ma, mw, cpl));
    } else if (one_row) OZ_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel<1>, *ma1, mw, cpl));
    else OZ_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel<2>, ma, mw, cpl));
    count_launch();
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

// OZ_CONV_STATIC=1 in the environment switches the convolution back to the static pass-to-cluster assignment (A/B runs).
bool conv_static_schedule() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("OZ_CONV_STATIC"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

void ResNet::free_workspaces() {
    for (Workspace &w : ws_) { cudaFree(w.x0); cudaFree(w.a); cudaFree(w.t); cudaFree(w.head_pre); cudaFree(w.work); }
    ws_.clear();
}

void launch_planes_to_nhwc32(const float *d_features, int planes, size_t n, size_t n_pad, void *d_out, cudaStream_t stream) {
    const size_t threads = n_pad * 256;
    planes_to_nhwc32_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_features, planes, n, n_pad, (__nv_bfloat16 *)d_out);
    count_launch();
}

ResNet::~ResNet() {
    cudaFree(d_blob_);
    free_workspaces();
}

int ResNet::load(const oz_resnet_shape &shape, const void *blob, size_t bytes) {
    if (shape.channels < 128 || shape.channels % 128 != 0 || shape.channels > 384)
        return fail(OZ_ERR_INVALID_ARGUMENT, "conv_channels must be 128, 256 or 384 for the native evaluator");
    if (shape.in_planes < 1 || shape.in_planes > 32)
        return fail(OZ_ERR_INVALID_ARGUMENT, "in_planes must be in [1, 32] (history_size <= 15)");
    if (shape.blocks < 0 || shape.value_hidden < 1 || shape.value_hidden > 512)
        return fail(OZ_ERR_INVALID_ARGUMENT, "unsupported residual block / value head size");
    ResNetLayout L = resnet_layout(shape);
    if (bytes != L.head_w) return fail(OZ_ERR_INVALID_ARGUMENT, "weight blob has " + std::to_string(bytes) + " bytes, expected " + std::to_string(L.head_w));
    const bool same = d_blob_ && shape.channels == shape_.channels && shape.blocks == shape_.blocks &&
                      shape.value_hidden == shape_.value_hidden && shape.in_planes == shape_.in_planes;
    if (!same) {
        cudaFree(d_blob_); d_blob_ = nullptr;
        OZ_CUDA(cudaMalloc(&d_blob_, L.total));
        map_w_.assign(L.conv_w.size(), CUtensorMap{});
        for (size_t i = 0; i < L.conv_w.size(); ++i)
            OZ_TRY(make_w_map(&map_w_[i], d_blob_ + L.conv_w[i], i == 0 ? 32 : shape.channels, shape.channels, 64));
        if (shape.channels != shape_.channels) free_workspaces();
        has_wr_ = false;
        if (shape.channels == 128 && shape.blocks >= 1 && L.conv_w.size() <= 64) {   // cluster tower (launch-bound batches)
            const size_t pitch = L.conv_w.size() > 2 ? L.conv_w[2] - L.conv_w[1] : 9 * 128 * 128 * 2 + 512;
            OZ_TRY(make_w_map_layers(&map_wr_, d_blob_ + L.conv_w[1], 128, (int)L.conv_w.size() - 1, pitch));
            has_wr_ = true;
        }
    }
    shape_ = shape; layout_ = L;
    OZ_CUDA(cudaMemcpy(d_blob_, blob, L.head_w, cudaMemcpyHostToDevice));
    OZ_CUDA(cudaMemcpy(d_blob_ + L.head_w, (const char *)blob + L.pconv_w, 2 * (size_t)shape.channels * 4, cudaMemcpyHostToDevice));
    OZ_CUDA(cudaMemcpy(d_blob_ + L.head_w + 2 * (size_t)shape.channels * 4, (const char *)blob + L.vconv_w, (size_t)shape.channels * 4, cudaMemcpyHostToDevice));
    return OZ_OK;
}

int ResNet::reserve(size_t boards, int ws) {
    if (ws < 0 || ws >= 16) return fail(OZ_ERR_INVALID_ARGUMENT, "workspace index out of range");
    if ((size_t)ws >= ws_.size()) ws_.resize((size_t)ws + 1);
    Workspace &w = ws_[(size_t)ws];
    const size_t pad = (boards + kBoardPad - 1) / kBoardPad * kBoardPad;
    if (pad <= w.cap) return OZ_OK;
    cudaFree(w.x0); cudaFree(w.a); cudaFree(w.t); cudaFree(w.head_pre); cudaFree(w.work);
    w = Workspace{};
    const size_t C = shape_.channels;
    OZ_CUDA(cudaMalloc(&w.x0, pad * 64 * 32 * 2));
    OZ_CUDA(cudaMalloc(&w.a, pad * 64 * C * 2));
    OZ_CUDA(cudaMalloc(&w.t, pad * 64 * C * 2));
    OZ_CUDA(cudaMalloc(&w.head_pre, pad * 64 * 4 * sizeof(float)));
    OZ_CUDA(cudaMalloc(&w.work, kMaxConvLayers * sizeof(int)));
    OZ_CUDA(cudaMemset(w.x0, 0, pad * 64 * 32 * 2));
    OZ_CUDA(cudaMemset(w.a, 0, pad * 64 * C * 2));
    OZ_CUDA(cudaMemset(w.t, 0, pad * 64 * C * 2));
    OZ_TRY(make_act_map(&w.map_x0, w.x0, pad, 32, pair::kBoxYp));
    OZ_TRY(make_act_map(&w.map_a, w.a, pad, (int)C, pair::kBoxYp));
    OZ_TRY(make_act_map(&w.map_t, w.t, pad, (int)C, pair::kBoxYp));
    OZ_TRY(make_act_map(&w.map_x0_1, w.x0, pad, 32, 3));          // one board row per pass (small batches)
    OZ_TRY(make_act_map(&w.map_a_1, w.a, pad, (int)C, 3));
    OZ_TRY(make_act_map(&w.map_t_1, w.t, pad, (int)C, 3));
    w.cap = pad;
    return OZ_OK;
}

int ResNet::input_buffer(size_t boards, int ws, void **x0) {
    if (!loaded()) return fail(OZ_ERR_STATE, "no network weights loaded");
    OZ_TRY(reserve(boards, ws));
    *x0 = ws_[(size_t)ws].x0;
    return OZ_OK;
}

int ResNet::forward(const float *d_features, size_t n, float *d_policy, float *d_value, cudaStream_t stream,
                    cudaEvent_t conv_begin, cudaEvent_t conv_end, int ws) {
    if (!loaded()) return fail(OZ_ERR_STATE, "no network weights loaded");
    if (n == 0) return OZ_OK;
    OZ_TRY(reserve(n, ws));
    const Workspace &w = ws_[(size_t)ws];
    dynamic_ = !conv_static_schedule();
    void *const d_x0_ = w.x0, *const d_a_ = w.a, *const d_t_ = w.t;
    float *const d_head_pre_ = w.head_pre;
    const CUtensorMap &map_x0_ = w.map_x0, &map_a_ = w.map_a, &map_t_ = w.map_t;
    const CUtensorMap *map_x0_1 = &w.map_x0_1, *map_a_1 = &w.map_a_1, *map_t_1 = &w.map_t_1;
    const size_t pad = (n + kBoardPad - 1) / kBoardPad * kBoardPad;
    const int C = shape_.channels;
    if (d_features) {                                   // nullptr: the caller staged the input in input_buffer()
        const size_t threads = pad * 256;
        planes_to_nhwc32_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_features, shape_.in_planes, n, pad, (__nv_bfloat16 *)d_x0_);
        count_launch();
    }
    // Last layer of a 128-channel tower (pair kernel): the epilogue also applies the three 1x1 head convolutions to its
    // fp32 values, so the final activation tensor is never written and the head kernel reads 1 KB instead of 16 KB per board.
    const bool fuse_heads = C == 128 && shape_.blocks > 0 && d_head_pre_ != nullptr;
    const int nconv = 1 + 2 * shape_.blocks;
    const bool use_tower = fuse_heads && has_wr_ && tower_enabled() && (int)(pad / 32) <= kTowerWaves * tower_max_clusters();
    if (fuse_heads && !use_tower) OZ_CUDA(cudaMemsetAsync(d_head_pre_, 0, pad * 64 * 4 * sizeof(float), stream));   // up front: nothing sits between the conv launches
    // One pass counter per convolution launch of this forward (dynamic pass scheduler), zeroed up front for the same reason.
    int *work = (dynamic_ && nconv <= kMaxConvLayers && !use_tower) ? w.work : nullptr;
    if (work) OZ_CUDA(cudaMemsetAsync(work, 0, (size_t)nconv * sizeof(int), stream));
    if (conv_begin) cudaEventRecord(conv_begin, stream);
    // Launch-bound batches of a 128-channel network: the whole tower in one cluster launch (tower_cluster_kernel).
    if (use_tower) {
        OZ_TRY(launch_tower(w, pad, stream));
        if (conv_end) cudaEventRecord(conv_end, stream);
        return launch_heads(w, n, fuse_heads, d_policy, d_value, stream, pdl_ && conv_end == nullptr);
    }
    ConvParams cp{};
    cp.units = (int)(pad / kBoardsPerUnit); cp.cout_total = C; cp.relu = 1;
#ifdef OZ_CONV_TIMELINE
    static unsigned long long *tl = nullptr;
    if (!tl) { cudaMalloc(&tl, 128 * 8 * sizeof(unsigned long long)); cudaMemset(tl, 0, 128 * 8 * sizeof(unsigned long long)); }
    g_conv_tl = tl;
    int tl_layer = 0;
    #define OZ_TL_NEXT() (cp.tl = tl + 8 * (tl_layer++))
#else
    #define OZ_TL_NEXT() ((void)0)
#endif
    // conv block (neural_net.py:9-29)
    cp.kblocks = 1; cp.skip = nullptr; cp.out = (__nv_bfloat16 *)d_a_; cp.bias = (const float *)(d_blob_ + layout_.conv_b[0]);
    cp.work = work;
    OZ_TL_NEXT();
    OZ_TRY(launch_conv(map_x0_, map_x0_1, map_w_[0], cp, stream, pdl_));
    cp.kblocks = C / kKBlock;
    for (int b = 0; b < shape_.blocks; ++b) {   // residual tower (neural_net.py:32-65)
        cp.skip = nullptr; cp.out = (__nv_bfloat16 *)d_t_; cp.bias = (const float *)(d_blob_ + layout_.conv_b[1 + 2 * b]);
        cp.work = work ? work + 1 + 2 * b : nullptr;
        OZ_TL_NEXT();
        OZ_TRY(launch_conv(map_a_, map_a_1, map_w_[1 + 2 * b], cp, stream, pdl_));
        cp.skip = (const __nv_bfloat16 *)d_a_; cp.out = (__nv_bfloat16 *)d_a_; cp.bias = (const float *)(d_blob_ + layout_.conv_b[2 + 2 * b]);
        cp.work = work ? work + 2 + 2 * b : nullptr;
        if (fuse_heads && b == shape_.blocks - 1) { cp.out = nullptr; cp.head_w = (const float *)(d_blob_ + layout_.head_w); cp.head_out = d_head_pre_; }
        OZ_TL_NEXT();
        OZ_TRY(launch_conv(map_t_, map_t_1, map_w_[2 + 2 * b], cp, stream, pdl_));
        cp.head_w = nullptr; cp.head_out = nullptr;
    }
    if (conv_end) cudaEventRecord(conv_end, stream);
    return launch_heads(w, n, fuse_heads, d_policy, d_value, stream);
}

int ResNet::launch_heads(const Workspace &w, size_t n, bool fuse_heads, float *d_policy, float *d_value, cudaStream_t stream, bool pdl) {
    const int C = shape_.channels;
    HeadParams hp{};
    hp.act = (const __nv_bfloat16 *)w.a;
    hp.pconv_w = (const float *)(d_blob_ + layout_.pconv_w); hp.pconv_b = (const float *)(d_blob_ + layout_.pconv_b);
    hp.pfc_w = (const float *)(d_blob_ + layout_.pfc_w); hp.pfc_b = (const float *)(d_blob_ + layout_.pfc_b);
    hp.vconv_w = (const float *)(d_blob_ + layout_.vconv_w); hp.vconv_b = (const float *)(d_blob_ + layout_.vconv_b);
    hp.vfc1_w = (const float *)(d_blob_ + layout_.vfc1_w); hp.vfc1_b = (const float *)(d_blob_ + layout_.vfc1_b);
    hp.vfc2_w = (const float *)(d_blob_ + layout_.vfc2_w); hp.vfc2_b = (const float *)(d_blob_ + layout_.vfc2_b);
    hp.pre = fuse_heads ? w.head_pre : nullptr;
    hp.policy = d_policy; hp.value = d_value; hp.C = C; hp.Hv = shape_.value_hidden; hp.n = n;
    // one board per warp while that still is a single wave of CTAs (2 per SM), four boards per warp beyond
    const int forced = heads_boards_override();
    const bool narrow = forced ? forced == 1 : n <= (size_t)148 * 2 * kHeadWarps;
    const int bpw = narrow ? 1 : kHeadBoardsWide;
    const size_t hsm = (size_t)heads_smem_floats(C, hp.Hv, bpw) * sizeof(float);
    static size_t hsm_set[2][64] = {};                   // per instantiation and device ordinal: the attribute belongs to the current device
    int hdev = 0;
    cudaGetDevice(&hdev);
    if (hsm > hsm_set[narrow][hdev & 63]) {
        if (narrow) OZ_CUDA(cudaFuncSetAttribute(heads_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
        else OZ_CUDA(cudaFuncSetAttribute(heads_kernel<kHeadBoardsWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
        hsm_set[narrow][hdev & 63] = hsm;
    }
    const size_t want = (n + kHeadWarps * bpw - 1) / (kHeadWarps * bpw);
    const unsigned grid = (unsigned)(want < 148 * 2 ? want : 148 * 2);
    if (narrow) OZ_CUDA(launch_kernel(heads_kernel<1>, dim3(grid), dim3(kHeadWarps * 32), hsm, stream, pdl, hp));
    else OZ_CUDA(launch_kernel(heads_kernel<kHeadBoardsWide>, dim3(grid), dim3(kHeadWarps * 32), hsm, stream, pdl, hp));
    count_launch();
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

int ResNet::conv3x3(const void *d_in, int cin, const void *d_w, const float *d_bias, const void *d_skip, int relu,
                    void *d_out, int cout, size_t boards_padded, cudaStream_t stream, int *d_work, float *d_stats, int stats_copies) {
    if (cin % kKBlock || cout % 128 || boards_padded % kBoardPad)
        return fail(OZ_ERR_INVALID_ARGUMENT, "conv3x3: cin % 32, cout % 128, boards % 32 must be 0");
    CUtensorMap ma, ma1, mw;
    OZ_TRY(make_act_map(&ma, const_cast<void *>(d_in), boards_padded, cin, pair::kBoxYp));
    OZ_TRY(make_act_map(&ma1, const_cast<void *>(d_in), boards_padded, cin, 3));
    OZ_TRY(make_w_map(&mw, const_cast<void *>(d_w), cin, cout, 64));
    ConvParams cp{};
    cp.skip = (const __nv_bfloat16 *)d_skip; cp.out = (__nv_bfloat16 *)d_out; cp.bias = d_bias;
    cp.units = (int)(boards_padded / kBoardsPerUnit); cp.kblocks = cin / kKBlock; cp.cout_total = cout; cp.relu = relu;
    cp.stats = d_stats; cp.stats_copies = stats_copies > 0 ? stats_copies : 1;
    if (d_work && !conv_static_schedule()) {          // dynamic pass scheduler: the caller's counter, zeroed in stream order
        OZ_CUDA(cudaMemsetAsync(d_work, 0, sizeof(int), stream));
        cp.work = d_work;
    }
    return launch_conv(ma, &ma1, mw, cp, stream);
}

}  // namespace oz
