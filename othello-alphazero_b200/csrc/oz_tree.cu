// Tree kernels: PUCT selection with virtual loss, feature encoding, expansion + backup,
// action choice, re-rooting.  One warp owns one game; lanes are children / path nodes /
// board squares.  Compiled with -fmad=false and explicit round-to-nearest intrinsics so
// that every float decision matches the reference's scalar code (search_thread.cpp:192-260
// compiled without FMA contraction) bit for bit.
#include <algorithm>
#include <cstdlib>
#include <cfloat>
#include <climits>

#include "oz_common.cuh"
#include "oz_rules.cuh"
#include "oz_tree.cuh"
#include "oz_tc.cuh"

namespace oz {

// ------------------------------------------------------------------ counter-based RNG
// Philox4x32-10 keyed by the engine seed; counter = (game id, search or ply, simulation, stream).
__device__ __forceinline__ void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// Gamma(alpha, 1) draw for child `child` of the root at simulation `sim`
// (reference: std::gamma_distribution, search_thread.cpp:23,233).
__device__ float gamma_draw(const Tree &t, uint32_t game_id, uint32_t search, uint32_t sim, int child) {
    uint32_t r[4];
    if (t.gamma_table) {
        philox(game_id, search, sim, 0x100u + (uint32_t)child, t.seed_lo, t.seed_hi, r);
        return t.gamma_table[r[0] % (uint32_t)t.gamma_len];
    }
    if (!(t.alpha > 0.0f)) return 0.0f;
    const float a = t.alpha < 1.0f ? t.alpha + 1.0f : t.alpha;   // Marsaglia-Tsang, boosted for alpha < 1
    const float d = a - 1.0f / 3.0f, c = 1.0f / sqrtf(9.0f * d);
    for (uint32_t attempt = 0;; ++attempt) {
        philox(game_id, search, sim, 0x100u + (uint32_t)child + 64u * attempt, t.seed_lo, t.seed_hi, r);
        const float u1 = u01(r[0]), u2 = u01(r[1]), u3 = u01(r[2]), u4 = u01(r[3]);
        const float x = sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
        float v = 1.0f + c * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        if (logf(u3) < 0.5f * x * x + d - d * v + d * logf(v) || attempt >= 15) {
            float g = d * v;
            if (t.alpha < 1.0f) g *= powf(u4, 1.0f / t.alpha);
            return g;
        }
    }
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Action index (0 = a1) of the j-th legal move in ascending action order (position.h:308-326).
__device__ __forceinline__ int nth_action(uint64_t legal, int j) {
    uint64_t rev = __brevll(legal);                  // bit a <-> action a
    for (int k = 0; k < j; ++k) rev &= rev - 1;
    return __ffsll((long long)rev) - 1;
}

__device__ __forceinline__ size_t arena_base(const Tree &t, int g) { return ((size_t)g * 2 + t.g_which[g]) * t.cap; }

// ------------------------------------------------------------------ reset (mcts.cpp:40-43)
__global__ void reset_kernel(Tree t) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= t.G) return;
    t.g_which[g] = 0;
    const size_t b = (size_t)g * 2 * t.cap;
    t.n_p1[b] = kInitialP1; t.n_p2[b] = kInitialP2;
    t.n_meta[b] = 1; t.n_parent[b] = -1; t.n_child[b] = -1;
    t.n_N[b] = 0; t.n_W[b] = 0.f; t.n_Q[b] = 0.f; t.n_P[b] = 1.f;
    t.g_root[g] = 0; t.g_count[g] = 1; t.g_status[g] = 0; t.g_ply[g] = 0; t.g_search[g] = 0; t.h_n[g] = 0;
    t.g_gameid[g] = t.first_game_id + (uint32_t)g;
    t.r_p1[g] = kInitialP1; t.r_p2[g] = kInitialP2; t.r_player[g] = 1;
}

// ------------------------------------------------------------------ select (search_thread.cpp:60-79, 192-260)
// GW warps of a block work on one game (the pipelined search launches single-warp blocks: 5 KB of static shared memory,
// they fit beside a resident convolution CTA).  The reference selects the B leaves of a round one after the other; what
// makes that order matter is the virtual loss, and the virtual loss only couples selections that meet at the SAME node:
//   * what a node decides for an arriving selection depends on the node's own children and on the selections that arrived
//     there earlier in slot order -- the rate-table index is the node's count before the arrival, the children's N / W / Q
//     change only by the virtual losses of earlier arrivals at this node, Dirichlet noise exists only at the root;
//   * so a chunk of 32 selections advances LEVEL BY LEVEL: the root's arrivals are decided first, in slot order, from
//     registers (the root's children stay there for the whole round); below the root the slots standing on the same node
//     are grouped, each distinct node's child block (N, W, Q, P and the header of every child) is read once, its arrivals
//     are decided in slot order from registers, the touched children are written back once;
//   * the virtual loss (N += 1, W -= 1, Q = W / N, search_thread.cpp:70-76) is applied by the lane that holds the chosen
//     child as the selection passes -- the reference applies it after the descent, which gives the same numbers because a
//     selection visits every node of its path once;
//   * everything a decision needs that does not depend on the earlier arrivals is computed beside the chain: the
//     Dirichlet draws and the noise weight of all 32 selections up front (lane = selection), the exploration factor
//     rate(N) * sqrt(sum of the children's counts) of every arrival (the sum grows by exactly one per arrival).
// Round 1 walked one selection at a time, a dependent chain of some 200 instructions and an L2 round trip per level:
// 105 us per round for a single game, now about 35.
// PUCT over the children a warp holds (lane = child, a 33rd child in the *1 registers): ucb = Q + (mult * P) / (1 + N),
// first maximum wins (search_thread.cpp:192-228).
__device__ __forceinline__ int puct_pick(int nch, float mult, int n0, int n1, float q0, float q1, float p0, float p1,
                                         bool v0, bool v1, int c0, int c1) {
    float u = -INFINITY;
    int idx = INT_MAX;
    if (v0) { u = __fadd_rn(q0, __fdiv_rn(__fmul_rn(mult, p0), __fadd_rn(1.0f, (float)n0))); idx = c0; }
    if (nch > 32 && v1) {                            // a 33rd child exists in a handful of positions only: warp-uniform branch
        const float u1 = __fadd_rn(q1, __fdiv_rn(__fmul_rn(mult, p1), __fadd_rn(1.0f, (float)n1)));
        if (u1 > u) { u = u1; idx = c1; }
    }
    // First maximum wins (strict > in the reference): the largest value through an order-preserving integer key
    // (-0 folded into +0 first, they compare equal as floats), then the smallest index that holds it.
    uint32_t key = __float_as_uint(__fadd_rn(u, 0.0f));
    key ^= (uint32_t)((int32_t)key >> 31) | 0x80000000u;
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
    return __reduce_min_sync(0xffffffffu, key == kmax ? idx : INT_MAX);
}
// Statistics and headers of all children of a node (first child at `a`), lane = child: one round trip.
__device__ __forceinline__ void load_child_block(const Tree &t, size_t a, int nch, int lane, int &n0, float &w0, float &q0, float &p0, int &cm0, int &cf0,
                                                 int &n1, float &w1, float &q1, float &p1, int &cm1, int &cf1) {
    const bool v0 = lane < nch, v1 = lane + 32 < nch;
    const size_t a0 = a + lane, a1 = a + lane + 32;
    n0 = v0 ? t.n_N[a0] : 0; w0 = v0 ? t.n_W[a0] : 0.f; q0 = v0 ? t.n_Q[a0] : 0.f; p0 = v0 ? t.n_P[a0] : 0.f;
    cm0 = v0 ? t.n_meta[a0] : 0; cf0 = v0 ? t.n_child[a0] : -1;
    n1 = 0; w1 = 0.f; q1 = 0.f; p1 = 0.f; cm1 = 0; cf1 = -1;
    if (nch > 32) {
        n1 = v1 ? t.n_N[a1] : 0; w1 = v1 ? t.n_W[a1] : 0.f; q1 = v1 ? t.n_Q[a1] : 0.f; p1 = v1 ? t.n_P[a1] : 0.f;
        cm1 = v1 ? t.n_meta[a1] : 0; cf1 = v1 ? t.n_child[a1] : -1;
    }
}

template <int GW>
__global__ void __launch_bounds__(GW * 32) select_kernel(Tree t) {
    __shared__ float s_gamma[32][33];                          // [selection of the chunk][root child]
    __shared__ float s_gsum[32];                               // eps / (sum of the selection's draws)
    // per-slot state of the chunk, handed from level to level (and between the warps of the game)
    __shared__ int sh_node[32], sh_m[32], sh_fc[32], sh_nn[32], sh_depth[32];
    __shared__ float sh_rate[32];
    griddep_launch_dependents();                               // launch-bound searches chain their kernels (launch_kernel, oz_common.cuh)
    griddep_wait();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int B = t.inflight, round = *t.round;
    unsigned long long path_nodes = 0, children_read = 0;      // traffic counters (oz_counters)
    for (int g = blockIdx.x; g < t.G; g += gridDim.x) {
    int *leaf_out = t.s_leaf + (size_t)g * B;
    if (t.g_status[g] & 1) {                         // finished game: nothing in flight
        for (int i = threadIdx.x; i < B; i += GW * 32) leaf_out[i] = -1;
        continue;
    }
    const size_t base = arena_base(t, g);
    const int root = t.g_root[g];
    const uint32_t game_id = t.g_gameid[g], search = t.g_search[g];
    // root header and root children, cached for the round (warp 0 decides at the root)
    const int rm = t.n_meta[base + root], rfc = t.n_child[base + root];
    int root_n = t.n_N[base + root];
    int rn0 = 0, rn1 = 0, rcm0 = 0, rcm1 = 0, rcf0 = -1, rcf1 = -1;
    float rw0 = 0.f, rw1 = 0.f, rq0 = 0.f, rq1 = 0.f, rp0 = 0.f, rp1 = 0.f;
    bool rdirty0 = false, rdirty1 = false;
    if (wib == 0 && (rm & 3) != 0) {
        const int rnch = rm >> 8;
        children_read += (unsigned)rnch;
        if (lane < rnch) { const size_t a = base + rfc + lane; rn0 = t.n_N[a]; rw0 = t.n_W[a]; rq0 = t.n_Q[a]; rp0 = t.n_P[a]; rcm0 = t.n_meta[a]; rcf0 = t.n_child[a]; }
        if (lane + 32 < rnch) { const size_t a = base + rfc + lane + 32; rn1 = t.n_N[a]; rw1 = t.n_W[a]; rq1 = t.n_Q[a]; rp1 = t.n_P[a]; rcm1 = t.n_meta[a]; rcf1 = t.n_child[a]; }
    }

    const bool noisy = t.eps > 0.0f && (rm & 3) != 0 && (rm >> 8) > 1;    // noise is only drawn where PUCT runs at the root
    const int rnch_all = rm >> 8;
    const bool root_open = (rm & 3) != 0 && rnch_all != 0;                // terminal or childless root: every selection ends there (:65)
    const float pm = __fsub_rn(1.0f, t.eps);
    for (int i0 = 0; i0 < B; i0 += 32) {
        const int chunk = min(32, B - i0);
        int *gpath_mine = t.s_path + ((size_t)g * B + i0 + (lane < chunk ? lane : 0)) * kMaxDepth;
        if (wib == 0) {
            if (noisy) {                                 // the chunk's Dirichlet draws, one selection per lane
                if (lane < chunk) {
                    float sum = 0.0f;                    // sequential float sum in child order (:236-241)
                    for (int c = 0; c < rnch_all; ++c) {
                        const float gc = gamma_draw(t, game_id, search, (uint32_t)(round * B + i0 + lane), c);
                        s_gamma[lane][c] = gc;
                        sum = __fadd_rn(sum, gc);
                    }
                    s_gsum[lane] = __fdiv_rn(t.eps, sum == 0.0f ? 1.0f : sum);   // the selection's noise weight eps / sum (:243-247), off the decision chain
                }
                __syncwarp();
            }
            // Per-slot state, lane = slot of the chunk: the node the selection stands on, its header (meta, first child), its
            // visit count BEFORE this selection's virtual loss (the rate-table index when the selection decides there) and
            // the exploration rate of that count, fetched by the slot's lane as soon as the count is known -- a level before
            // it is used (in the arrival loop it would be a dependent global load in front of every decision).
            int s_node = root, s_m = rm, s_fc = rfc, s_nn = 0, s_depth = 0;
            float s_rate = 0.f;
            // The exploration factor rate(N) * sqrt(sum of the children's counts) of the root's s-th arrival of the chunk is known
            // up front -- every arrival adds exactly one visit to the root and to one child -- so lane s computes it for its
            // selection (table load, IEEE square root) and the sequential loop below only fetches it with a shuffle.
            float mult_mine = 0.f;
            if (root_open && rnch_all > 1) {
                const int total0 = __reduce_add_sync(0xffffffffu, rn0 + rn1);
                int nn = root_n + lane;
                if (lane < chunk && nn >= t.rate_len) atomicOr(t.error, kErrRateTable);
                if (nn >= t.rate_len) nn = t.rate_len - 1;
                mult_mine = __fmul_rn(t.rate[nn], __fsqrt_rn((float)(total0 + lane)));   // rate: logf((1 + N + c_base) / c_base) + c_init, host libm
            }
            // ---- level 0: the root's arrivals in slot order, statistics in registers
            for (int s = 0; s < chunk; ++s) {
                int next = root, nm = rm, nfc = rfc, nnn = root_n, dep = 0;
                if (root_open) {
                    const int nch = rnch_all;
                    const int c0 = lane, c1 = lane + 32;
                    const bool v0 = c0 < nch, v1 = c1 < nch;
                    int idx = 0;                         // a single child is taken without any arithmetic (:194-196)
                    if (nch > 1) {
                        const float mult = __shfl_sync(0xffffffffu, mult_mine, s);
                        float p0 = rp0, p1 = rp1;
                        if (noisy) {                     // Dirichlet noise, re-drawn per selection (:230-249)
                            const float g0 = v0 ? s_gamma[s][c0] : 0.f;
                            const float g1 = (nch > 32 && v1) ? s_gamma[s][c1] : 0.f;
                            const float nm2 = s_gsum[s];
                            p0 = __fadd_rn(__fmul_rn(p0, pm), __fmul_rn(g0, nm2));
                            if (nch > 32) p1 = __fadd_rn(__fmul_rn(p1, pm), __fmul_rn(g1, nm2));
                        }
                        idx = puct_pick(nch, mult, rn0, rn1, rq0, rq1, p0, p1, v0, v1, c0, c1);
                    }
                    next = rfc + idx;
                    const bool hi = idx >= 32;
                    nm = __shfl_sync(0xffffffffu, hi ? rcm1 : rcm0, idx & 31);
                    nfc = __shfl_sync(0xffffffffu, hi ? rcf1 : rcf0, idx & 31);
                    nnn = __shfl_sync(0xffffffffu, hi ? rn1 : rn0, idx & 31);     // the child's count BEFORE this selection's virtual loss
                    if (lane == (idx & 31)) {            // virtual loss by the lane that holds the chosen child (written back after the chunk)
                        if (hi) { rn1 += 1; rw1 = __fsub_rn(rw1, 1.0f); rq1 = __fdiv_rn(rw1, (float)rn1); rdirty1 = true; }
                        else { rn0 += 1; rw0 = __fsub_rn(rw0, 1.0f); rq0 = __fdiv_rn(rw0, (float)rn0); rdirty0 = true; }
                    }
                    dep = 1;
                }
                root_n += 1;                             // root visit count (:78)
                if (lane == s) { s_node = next; s_m = nm; s_fc = nfc; s_nn = nnn; s_depth = dep; s_rate = t.rate[nnn < t.rate_len ? nnn : t.rate_len - 1]; }
            }
            if (lane < chunk && s_depth == 1) gpath_mine[0] = s_node;
            sh_node[lane] = s_node; sh_m[lane] = s_m; sh_fc[lane] = s_fc; sh_nn[lane] = s_nn; sh_depth[lane] = s_depth; sh_rate[lane] = s_rate;
        }
        // ---- levels >= 1.  What a node decides for an arriving selection depends on the node's own children and on the
        // selections that arrived there EARLIER in slot order -- on nothing else in the tree.  So the chunk advances level by
        // level: the slots standing on the same node are grouped (__match_any_sync), each distinct node's child block is read
        // ONCE (the next node's block is already in flight while the current node's arrivals are decided), its arrivals are
        // decided in slot order from registers, and the touched children are written back once.  Distinct nodes of a level
        // are independent of each other: the GW warps of the block take them in turn (the slot state travels through shared
        // memory), so a level costs its busiest warp's decisions, not the sum -- a decision is some 200 dependent
        // instructions (IEEE square root, two IEEE divisions, three warp reductions), about 950 clocks, and there are 120 to
        // 150 of them below the root in a 32-selection round.
        for (int level = 1;; ++level) {
            __syncthreads();                             // the previous level's (or the root pass's) slot state is in shared memory
            int s_node = sh_node[lane], s_m = sh_m[lane], s_fc = sh_fc[lane], s_nn = sh_nn[lane], s_depth = sh_depth[lane];
            float s_rate = sh_rate[lane];
            const bool active = lane < chunk && s_depth == level && (s_m & 3) != 0 && (s_m >> 8) != 0;
            if (!__ballot_sync(0xffffffffu, active)) break;       // the same for every warp of the block
            const unsigned grp = __match_any_sync(0xffffffffu, active ? s_node : -1 - lane);
            const unsigned all_leaders = __ballot_sync(0xffffffffu, active && (__ffs(grp) - 1) == lane);
            // this warp's share: the leaders whose rank among the level's leaders is wib modulo GW
            unsigned leaders = 0;
            {
                unsigned rest = all_leaders;
                for (int r = 0; rest; ++r) { const unsigned low = rest & (0u - rest); if (r % GW == wib) leaders |= low; rest ^= low; }
            }
            __syncthreads();                             // everybody has read the state before anybody writes the next level's
            unsigned touched = 0;                        // slots whose state this warp advanced
            if (leaders) {
                int L = __ffs(leaders) - 1;
                leaders &= leaders - 1;
                int fcX = __shfl_sync(0xffffffffu, s_fc, L), nchX = __shfl_sync(0xffffffffu, s_m, L) >> 8;
                unsigned arr = __shfl_sync(0xffffffffu, grp, L);
                int n0, n1, cm0, cm1, cf0, cf1;
                float w0, w1, q0, q1, p0, p1;
                load_child_block(t, base + fcX, nchX, lane, n0, w0, q0, p0, cm0, cf0, n1, w1, q1, p1, cm1, cf1);
                while (true) {
                    // the next distinct node's block: in flight while this node's arrivals are decided
                    const bool more = leaders != 0;
                    int fcY = 0, nchY = 0;
                    unsigned arrY = 0;
                    int yn0 = 0, yn1 = 0, ycm0 = 0, ycm1 = 0, ycf0 = -1, ycf1 = -1;
                    float yw0 = 0.f, yw1 = 0.f, yq0 = 0.f, yq1 = 0.f, yp0 = 0.f, yp1 = 0.f;
                    if (more) {
                        const int L2 = __ffs(leaders) - 1;
                        leaders &= leaders - 1;
                        fcY = __shfl_sync(0xffffffffu, s_fc, L2); nchY = __shfl_sync(0xffffffffu, s_m, L2) >> 8;
                        arrY = __shfl_sync(0xffffffffu, grp, L2);
                        load_child_block(t, base + fcY, nchY, lane, yn0, yw0, yq0, yp0, ycm0, ycf0, yn1, yw1, yq1, yp1, ycm1, ycf1);
                    }
                    // this node's arrivals, in slot order
                    const int nch = nchX;
                    const int c0 = lane, c1 = lane + 32;
                    const bool v0 = c0 < nch, v1 = c1 < nch;
                    bool dirty0 = false, dirty1 = false;
                    touched |= arr;
                    // the j-th arrival sees the children's counts sum to (sum on arrival of the first) + j: lane j takes its square root now
                    float sq_mine = 0.f;
                    if (nch > 1) sq_mine = __fsqrt_rn((float)(__reduce_add_sync(0xffffffffu, n0 + n1) + lane));
                    for (int j = 0; arr; ++j) {
                        const int s = __ffs(arr) - 1;
                        arr &= arr - 1;
                        int idx = 0;
                        if (nch > 1) {
                            const int nn = __shfl_sync(0xffffffffu, s_nn, s);
                            if (nn >= t.rate_len && lane == 0) atomicOr(t.error, kErrRateTable);
                            const float rate = __shfl_sync(0xffffffffu, s_rate, s);   // rate[nn]: logf((1 + N + c_base) / c_base) + c_init, host libm
                            children_read += (unsigned)nch;
                            const float mult = __fmul_rn(rate, __shfl_sync(0xffffffffu, sq_mine, j));
                            idx = puct_pick(nch, mult, n0, n1, q0, q1, p0, p1, v0, v1, c0, c1);
                        }
                        const int next = fcX + idx;
                        const bool hi = idx >= 32;
                        const int nm = __shfl_sync(0xffffffffu, hi ? cm1 : cm0, idx & 31);
                        const int nfc = __shfl_sync(0xffffffffu, hi ? cf1 : cf0, idx & 31);
                        const int nnn = __shfl_sync(0xffffffffu, hi ? n1 : n0, idx & 31);
                        if (lane == (idx & 31)) {        // virtual loss (search_thread.cpp:70-76), in the register copy
                            if (hi) { n1 += 1; w1 = __fsub_rn(w1, 1.0f); q1 = __fdiv_rn(w1, (float)n1); dirty1 = true; }
                            else { n0 += 1; w0 = __fsub_rn(w0, 1.0f); q0 = __fdiv_rn(w0, (float)n0); dirty0 = true; }
                        }
                        if (lane == s) {
                            if (s_depth < kMaxDepth) gpath_mine[s_depth] = next; else atomicOr(t.error, kErrDepth);
                            s_node = next; s_m = nm; s_fc = nfc; s_nn = nnn; s_depth += 1;
                            s_rate = t.rate[nnn < t.rate_len ? nnn : t.rate_len - 1];
                        }
                    }
                    if (dirty0) { const size_t a = base + fcX + c0; t.n_N[a] = n0; t.n_W[a] = w0; t.n_Q[a] = q0; }
                    if (dirty1) { const size_t a = base + fcX + c1; t.n_N[a] = n1; t.n_W[a] = w1; t.n_Q[a] = q1; }
                    if (!more) break;
                    fcX = fcY; nchX = nchY; arr = arrY;
                    n0 = yn0; n1 = yn1; cm0 = ycm0; cm1 = ycm1; cf0 = ycf0; cf1 = ycf1;
                    w0 = yw0; w1 = yw1; q0 = yq0; q1 = yq1; p0 = yp0; p1 = yp1;
                }
            }
            if (touched & (1u << lane)) {                // hand the advanced slots to the next level
                sh_node[lane] = s_node; sh_m[lane] = s_m; sh_fc[lane] = s_fc; sh_nn[lane] = s_nn; sh_depth[lane] = s_depth; sh_rate[lane] = s_rate;
            }
        }
        // the chunk's results (the loop was left behind a barrier: the state in shared memory is final)
        if (wib == 0) {
            const int d = sh_depth[lane];
            const int dclip = d > kMaxDepth ? kMaxDepth : d;
            path_nodes += (unsigned)__reduce_add_sync(0xffffffffu, lane < chunk ? dclip : 0);
            if (lane < chunk) { leaf_out[i0 + lane] = sh_node[lane]; t.s_len[(size_t)g * B + i0 + lane] = dclip; }
            // the root's children: written back once per chunk (later chunks read these registers, later kernels the memory)
            if (rdirty0) { const size_t a = base + rfc + lane; t.n_N[a] = rn0; t.n_W[a] = rw0; t.n_Q[a] = rq0; rdirty0 = false; }
            if (rdirty1) { const size_t a = base + rfc + lane + 32; t.n_N[a] = rn1; t.n_W[a] = rw1; t.n_Q[a] = rq1; rdirty1 = false; }
        }
        __syncthreads();                                 // this chunk's stores and shared state are done with before the next chunk
    }
    if (wib == 0 && lane == 0) t.n_N[base + root] = root_n;
    }
    if (lane == 0 && (path_nodes | children_read)) { unsigned long long *st = t.stats + 4 * ((blockIdx.x * GW + wib) % kStatCopies); atomicAdd(&st[1], path_nodes); atomicAdd(&st[2], children_read); }
}

// ------------------------------------------------------------------ encode (search_thread.cpp:87-100, transformation.h:83-116)
// Positions of the leaf of `slot` and its ancestors, newest first: the chain inside the arena, then the retired
// ancestors of the root (position_iterator.h:49-52 crosses the root the same way).  The chain is not walked through the
// parent links (one dependent read per ply of history): the selection stored its path, which IS the chain -- entry
// len - 1 is the leaf, entry 0 the root's child -- so lane k reads the k-th position back directly.
__device__ __forceinline__ int gather_history(const Tree &t, int g, size_t base, size_t slot, int lane, uint64_t *sp1, uint64_t *sp2) {
    const int len = t.s_len[slot], hn = t.h_n[g];
    const int n = min(t.H, len + 1 + hn);
    if (lane < n) {
        if (lane <= len) {
            const int nd = lane < len ? t.s_path[slot * kMaxDepth + (len - 1 - lane)] : t.g_root[g];
            sp1[lane] = t.n_p1[base + nd]; sp2[lane] = t.n_p2[base + nd];
        } else {
            const int r = lane - len - 1;
            sp1[lane] = t.h_p1[(size_t)g * kMaxHist + r]; sp2[lane] = t.h_p2[(size_t)g * kMaxHist + r];
        }
    }
    __syncwarp();
    return n;
}

__device__ __forceinline__ void write_planes(float *f, int H, int nh, const uint64_t *sp1, const uint64_t *sp2,
                                             float side, int tf, int lane) {
    f[lane] = side; f[lane + 32] = side;
    const int i0 = lane, i1 = lane + 32;
    const int j0 = transform_action(i0, tf), j1 = transform_action(i1, tf);
    const uint64_t m0 = 1ull << (63 - i0), m1 = 1ull << (63 - i1);
    for (int k = 0; k < H; ++k) {
        float *pl = f + 64 + k * 128;
        if (k >= nh) { pl[lane] = 0.f; pl[lane + 32] = 0.f; pl[64 + lane] = 0.f; pl[96 + lane] = 0.f; continue; }
        const uint64_t b = sp1[k], w = sp2[k];
        pl[j0] = (b & m0) ? 1.f : 0.f; pl[j1] = (b & m1) ? 1.f : 0.f;
        pl[64 + j0] = (w & m0) ? 1.f : 0.f; pl[64 + j1] = (w & m1) ? 1.f : 0.f;
    }
}

// The same planes written straight into the native network's input tensor: bf16 NHWC [slot][64][32], channel = plane
// (planes >= 1 + 2H stay zero).  Square i of the board lands on position transform_action(i, tf) (transformation.h:104-112).  To keep the stores coalesced a lane owns the TARGET positions lane and lane + 32
// and looks its source squares up through the inverse symmetry (a mirrored symmetry is its own inverse, a pure rotation
// by r quarter turns is undone by 4 - r).
__device__ __forceinline__ void write_planes_nhwc(uint16_t *x, int nh, const uint64_t *sp1, const uint64_t *sp2, int player, int tf, int lane) {
    const int inv = (tf & 1) ? tf : (((4 - (tf >> 1)) & 3) << 1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h, i = transform_action(j, inv);
        const uint64_t m = 1ull << (63 - i);
        uint32_t w[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) w[q] = 0u;
        if (player == 2) w[0] = 0x3F80u;                                   // plane 0 = player - 1 (transformation.h:91)
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            if (k < nh) {
                if (sp1[k] & m) w[(1 + 2 * k) >> 1] |= 0x3F800000u;       // odd channel 1 + 2k: high half
                if (sp2[k] & m) w[(2 + 2 * k) >> 1] |= 0x00003F80u;       // even channel 2 + 2k: low half
            }
        }
        uint4 *dst = (uint4 *)(x + (size_t)j * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) encode_kernel(Tree t, float *features, uint16_t *nhwc) {
    __shared__ uint64_t s1[WARPS][kMaxHist], s2[WARPS][kMaxHist];
    griddep_launch_dependents();
    griddep_wait();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t slots = (size_t)t.G * t.inflight;
    for (size_t slot = (size_t)blockIdx.x * WARPS + wib; slot < slots; slot += (size_t)gridDim.x * WARPS) {
        const int g = (int)(slot / t.inflight), i = (int)(slot % t.inflight);
        const int leaf = t.s_leaf[slot];
        if (leaf < 0) continue;
        const size_t base = arena_base(t, g);
        const int player = t.n_meta[base + leaf] & 3;
        if (player == 0) continue;                        // terminal leaf: slot keeps its stale planes (:88-90)
        int tf = 0;
        if (t.sym_mode == OZ_SYM_RANDOM) {
            uint32_t r[4];
            philox(t.g_gameid[g], t.g_search[g], (uint32_t)(*t.round * t.inflight + i), 0u, t.seed_lo, t.seed_hi, r);
            tf = (int)(r[0] & 7u);
        }
        if (lane == 0) t.s_tf[slot] = tf;
        const int nh = gather_history(t, g, base, slot, lane, s1[wib], s2[wib]);
        if (nhwc) write_planes_nhwc(nhwc + slot * 64 * 32, nh, s1[wib], s2[wib], player, tf, lane);
        else write_planes(features + slot * (size_t)(1 + 2 * t.H) * 64, t.H, nh, s1[wib], s2[wib], (float)player - 1.0f, tf, lane);
        __syncwarp();                                     // the history buffer is reused by the next slot
    }
}

// ------------------------------------------------------------------ expand + backup (search_thread.cpp:130-190)
// The reference handles the slots of a batch one after another: expand the leaf if it is still childless, then back its
// value up.  Backups must stay in slot order (they add floats to shared ancestors), expansions need not: creating the
// children of a leaf touches nothing a backup reads or writes.  So a chunk of 32 slots is processed as
//   1. one slot per lane: leaf, path length, symmetry, value, the leaf's header and discs (two memory round trips for the
//      chunk), which slots expand (non-terminal, childless, FIRST slot of the chunk holding that leaf -- later copies skip
//      the expansion, search_thread.cpp:135), the legal moves and child count of every expanding leaf (one SIMT pass of
//      the move generator for 32 leaves instead of 32 passes), a prefix sum that gives every leaf the node ids the
//      sequential order would have given it;
//   2. all new children of the chunk as one flat range, 32 per step with every lane busy: position after the move
//      (flips + the mover's next legal moves), prior gathered from the policy row through the symmetry;
//   3. the 32 backups in slot order.  Slot order only matters PER NODE (a node's W is a float sum in the order of the slots
//      whose paths cross it), and two path entries can be the same node only at the same depth.  So for paths of at most 16
//      nodes (practically always) the chunk's W values live in shared memory for the duration of the backups: lane = slot
//      finds, depth by depth, the first slot of the chunk holding the same node (__match_any_sync) and loads W and N of the
//      nodes it is first for (every load of the chunk in flight at once); then lane = depth k adds the 32 slots' values
//      into its depth's nodes in slot order -- 16 independent register/shared-memory chains, no memory round trip between
//      two backups (each backup used to wait for an L2 round trip: 32 x 0.45 us per round for a single game); then the
//      owners write W and Q = W / N back once.  Same float additions in the same order: the same bits.
// GW warps work on one game.  Creating children (2.) and backing up (3.) touch different things -- new nodes and the expanding
// leaves' headers on one side, W and Q of path nodes on the other -- so with GW > 1 warp 0 runs the backups while the other
// warps create the chunk's children, 32 * (GW - 1) at a step; every warp derives step 1 for itself (the same loads, the same
// numbers).  GW = 1 (the narrow kernel of the pipelined search) does 1, 2, 3 in turn.
constexpr int kFastBackupDepth = 16;
template <int GW>
__global__ void __launch_bounds__(GW * 32) expand_kernel(Tree t, const float *__restrict__ policy, const float *__restrict__ value) {
    __shared__ float s_bw[1][32 * kFastBackupDepth];          // W of the chunk's nodes, at [first slot][depth]
    __shared__ int s_bn[1][32 * kFastBackupDepth];            // their visit counts
    __shared__ __align__(16) uint8_t s_bfirst[1][32 * kFastBackupDepth];   // [slot][depth] -> first slot with that node
    __shared__ float s_bv[1][32];                             // value backed up from slot li
    __shared__ int s_blen[1][32];
    constexpr int WARPS = GW > 1 ? 4 : 1;                     // (the backup fast path below is compiled in for GW > 1)
    const int lane = threadIdx.x & 31, wib = 0, gwarp = threadIdx.x >> 5;
    const int B = t.inflight;
    unsigned long long expanded = 0;
    griddep_launch_dependents();
    griddep_wait();                                             // nothing of the tree or the round counter is touched before this
    if (blockIdx.x == 0 && threadIdx.x == 0) *t.round += 1;     // select / encode of this round are done
    for (int g = blockIdx.x; g < t.G; g += gridDim.x) {
    if (t.g_status[g] & 1) continue;
    const size_t base = arena_base(t, g);
    int count = t.g_count[g];                            // nodes in the arena; kept in a register, written back once
    for (int i0 = 0; i0 < B; i0 += 32) {
        const int chunk = min(32, B - i0);
        const size_t slot0 = (size_t)g * B + i0;
        // ---- 1. per-slot inputs, one slot per lane
        __syncwarp();
        const bool c_valid = lane < chunk;
        int c_leaf = -1 - lane, c_len = 0, c_tf = 0, c_meta = 0;
        float c_val = 0.f;
        uint64_t c_p1 = 0, c_p2 = 0;
        if (c_valid) {
            const size_t sl = slot0 + lane;
            c_leaf = t.s_leaf[sl]; c_len = t.s_len[sl]; c_tf = t.s_tf[sl]; c_val = value[sl];
            c_meta = t.n_meta[base + c_leaf]; c_p1 = t.n_p1[base + c_leaf]; c_p2 = t.n_p2[base + c_leaf];
        }
        if (GW > 1) __syncthreads();                      // every warp has read the leaves' headers before a creator rewrites one
        const int c_player = c_meta & 3;
        const unsigned same = __match_any_sync(0xffffffffu, c_leaf);          // lanes holding the same leaf
        const bool expands = c_valid && c_player != 0 && (c_meta >> 8) == 0 && (__ffs(same) - 1) == lane;
        const uint64_t c_me = c_player == 1 ? c_p1 : c_p2, c_opp = c_player == 1 ? c_p2 : c_p1;
        const uint64_t c_legal = expands ? legal_moves(c_me, c_opp) : 0;
        const uint64_t c_next = (expands && !c_legal) ? legal_moves(c_opp, c_me) : 0;
        const int c_n = expands ? (c_legal ? __popcll(c_legal) : 1) : 0;
        int incl = c_n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int total = __shfl_sync(0xffffffffu, incl, 31);
        const int c_off = incl - c_n;
        if (count + total > t.cap) {                     // cannot happen after oz_engine's capacity check; reported, nothing expanded
            if (lane == 0) atomicOr(t.error, kErrArenaFull);
            total = 0;
        }
        // ---- 2. the children of the chunk, flat (GW > 1: warps 1 .. GW - 1, warp 0 goes straight to the backups)
        for (int k0 = GW > 1 ? 32 * (gwarp - 1) : 0; k0 < total && (GW == 1 || gwarp > 0); k0 += GW > 1 ? 32 * (GW - 1) : 32) {
            const int k = k0 + lane;
            int sidx = 0;                                // first lane whose inclusive prefix exceeds k
#pragma unroll
            for (int step = 16; step; step >>= 1) {
                const int cand = sidx + step;
                const int v = __shfl_sync(0xffffffffu, incl, (cand - 1) & 31);
                if (cand <= 31 && v <= k) sidx = cand;
            }
            Position sp;
            sp.player = __shfl_sync(0xffffffffu, c_player, sidx);
            sp.p1 = __shfl_sync(0xffffffffu, c_p1, sidx); sp.p2 = __shfl_sync(0xffffffffu, c_p2, sidx);
            sp.legal = __shfl_sync(0xffffffffu, c_legal, sidx); sp.next_legal = __shfl_sync(0xffffffffu, c_next, sidx);
            const int leaf_s = __shfl_sync(0xffffffffu, c_leaf, sidx), tf_s = __shfl_sync(0xffffffffu, c_tf, sidx);
            const int j = k - __shfl_sync(0xffffffffu, c_off, sidx);
            if (k < total) {
                const int a = sp.legal ? nth_action(sp.legal, j) : 64;
                const Position c = apply_action(sp, a);
                const size_t d = base + count + k;
                t.n_p1[d] = c.p1; t.n_p2[d] = c.p2; t.n_meta[d] = c.player;
                t.n_parent[d] = leaf_s; t.n_child[d] = -1;
                t.n_N[d] = 0; t.n_W[d] = 0.f; t.n_Q[d] = 0.f;
                t.n_P[d] = policy[(slot0 + sidx) * 65 + transform_action(a, tf_s)];   // raw softmax output, no renormalisation (:139-149)
            }
        }
        if ((GW == 1 || gwarp == 1) && expands && total > 0) { t.n_child[base + c_leaf] = count + c_off; t.n_meta[base + c_leaf] = c_player | (c_n << 8); }
        count += total;
        if (gwarp == 0) expanded += (unsigned)total;
        __syncwarp();
        if (GW > 1 && gwarp > 0) { __syncthreads(); continue; }        // the creators are done with this chunk
        // ---- 3. backups in slot order.
        if (WARPS == 4 && __reduce_max_sync(0xffffffffu, c_valid ? c_len : 0) <= kFastBackupDepth) {
            const int bw = WARPS == 4 ? wib : 0;
            // value of every slot, lane = slot
            float my_v = 0.f;
            if (c_valid) {
                if (c_player != 0) my_v = -c_val;          // edge value is from the parent's view (:156-157)
                else {                                     // terminal: disc count from the parent's side (:158-177)
                    const int par = t.n_parent[base + c_leaf];
                    const int pp = par >= 0 ? (t.n_meta[base + par] & 3) : t.h_player[(size_t)g * kMaxHist];
                    const int pc = __popcll(pp == 1 ? c_p1 : c_p2), oc = __popcll(pp == 1 ? c_p2 : c_p1);
                    my_v = pc > oc ? 1.0f : (pc < oc ? -1.0f : 0.0f);
                }
            }
            int pnode[kFastBackupDepth];
            {
                const int4 *row = (const int4 *)(t.s_path + (slot0 + (c_valid ? lane : 0)) * kMaxDepth);
#pragma unroll
                for (int q = 0; q < kFastBackupDepth / 4; ++q) {
                    int4 v4 = make_int4(0, 0, 0, 0);
                    if (c_valid && 4 * q < c_len) v4 = row[q];
                    pnode[4 * q] = v4.x; pnode[4 * q + 1] = v4.y; pnode[4 * q + 2] = v4.z; pnode[4 * q + 3] = v4.w;
                }
            }
            uint32_t fpack[kFastBackupDepth / 4];
            unsigned mine = 0;                             // depths at which this slot is the first holder of its node
#pragma unroll
            for (int q = 0; q < kFastBackupDepth / 4; ++q) fpack[q] = 0;
#pragma unroll
            for (int k = 0; k < kFastBackupDepth; ++k) {
                const bool has = c_valid && k < c_len;
                const unsigned m = __match_any_sync(0xffffffffu, has ? pnode[k] : -1 - lane);
                const int f = __ffs(m) - 1;
                fpack[k >> 2] |= (uint32_t)f << (8 * (k & 3));
                if (has && f == lane) {
                    mine |= 1u << k;
                    s_bw[bw][lane * kFastBackupDepth + k] = t.n_W[base + pnode[k]];
                    s_bn[bw][lane * kFastBackupDepth + k] = t.n_N[base + pnode[k]];
                }
            }
            *(uint4 *)(s_bfirst[bw] + lane * kFastBackupDepth) = make_uint4(fpack[0], fpack[1], fpack[2], fpack[3]);
            s_bv[bw][lane] = my_v;
            s_blen[bw][lane] = c_valid ? c_len : 0;
            __syncwarp();
            if (lane < kFastBackupDepth) {                 // lane = depth: this depth's nodes, slots in order
                for (int li = 0; li < chunk; ++li) {
                    const int len = s_blen[bw][li];
                    if (lane < len) {
                        const int idx = (int)s_bfirst[bw][li * kFastBackupDepth + lane] * kFastBackupDepth + lane;
                        const float v = s_bv[bw][li];
                        const float vk = ((len - 1 - lane) & 1) ? -v : v;
                        s_bw[bw][idx] = __fadd_rn(s_bw[bw][idx], __fadd_rn(1.0f, vk));   // W += 1 + v cancels the virtual loss (:184)
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < kFastBackupDepth; ++k)
                if (mine & (1u << k)) {
                    const float w = s_bw[bw][lane * kFastBackupDepth + k];
                    const size_t a = base + pnode[k];
                    t.n_W[a] = w; t.n_Q[a] = __fdiv_rn(w, (float)s_bn[bw][lane * kFastBackupDepth + k]);
                }
            __syncwarp();
            if (GW > 1) __syncthreads();                  // chunk barrier (matches the creators')
            continue;
        }
        // Longer paths: one backup after the other.  The first eight path entries of every slot are fetched up front, transposed over
        // the warp (lane 8 * (s % 4) + k holds entry k of slots s = 4 r + ..., r = 0..7), so the sequential part below
        // starts with the nodes' W and N instead of a dependent read of the path; longer paths read the rest from memory.
        int pa[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int sl = 4 * r + (lane >> 3), k = lane & 7;
            const int ln = __shfl_sync(0xffffffffu, c_len, sl);
            pa[r] = (sl < chunk && k < ln) ? t.s_path[(slot0 + sl) * kMaxDepth + k] : -1;
        }
        for (int li = 0; li < chunk; ++li) {
            const int leaf = __shfl_sync(0xffffffffu, c_leaf, li);
            const int player = __shfl_sync(0xffffffffu, c_player, li);
            float v;
            if (player != 0) {
                v = -__shfl_sync(0xffffffffu, c_val, li);  // edge value is from the parent's view (:156-157)
            } else {                                       // terminal: disc count from the parent's side (:158-177)
                const uint64_t lp1 = __shfl_sync(0xffffffffu, c_p1, li), lp2 = __shfl_sync(0xffffffffu, c_p2, li);
                const int par = t.n_parent[base + leaf];
                const int pp = par >= 0 ? (t.n_meta[base + par] & 3) : t.h_player[(size_t)g * kMaxHist];
                const int pc = __popcll(pp == 1 ? lp1 : lp2), oc = __popcll(pp == 1 ? lp2 : lp1);
                v = pc > oc ? 1.0f : (pc < oc ? -1.0f : 0.0f);
            }
            // along the stored path (leaf last); the sign alternates towards the root
            const int len = __shfl_sync(0xffffffffu, c_len, li);
            const int *gpath = t.s_path + (slot0 + li) * kMaxDepth;
            int par_r = pa[0];
#pragma unroll
            for (int r = 1; r < 8; ++r) if ((li >> 2) == r) par_r = pa[r];
            const int pre = __shfl_sync(0xffffffffu, par_r, ((li & 3) << 3) + (lane & 7));
            for (int k = lane; k < len; k += 32) {
                const size_t a = base + (k < 8 ? pre : gpath[k]);
                const float vk = ((len - 1 - k) & 1) ? -v : v;
                const float w = __fadd_rn(t.n_W[a], __fadd_rn(1.0f, vk));   // W += 1 + v cancels the virtual loss (:184)
                t.n_W[a] = w; t.n_Q[a] = __fdiv_rn(w, (float)t.n_N[a]);
            }
            __syncwarp();
        }
        if (GW > 1) __syncthreads();                      // chunk barrier (matches the creators')
    }
    if (gwarp == 0 && lane == 0) t.g_count[g] = count;
    }
    if (gwarp == 0 && lane == 0 && expanded) atomicAdd(&t.stats[4 * (blockIdx.x % kStatCopies)], expanded);
}

__global__ void end_search_kernel(Tree t) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < t.G) t.g_search[g] += 1;
}

// ------------------------------------------------------------------ root statistics (mcts.cpp:45-61)
__global__ void root_stats_kernel(Tree t, int32_t *n_children, int32_t *counts, float *q, int32_t *root_n) {
    const int lane = threadIdx.x & 31;
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= t.G) return;
    const size_t base = arena_base(t, g);
    const int root = t.g_root[g];
    const int nch = t.n_meta[base + root] >> 8, fc = t.n_child[base + root];
    if (lane == 0) { if (n_children) n_children[g] = nch; if (root_n) root_n[g] = t.n_N[base + root]; }
    for (int j = lane; j < kStride; j += 32) {
        if (counts) counts[(size_t)g * kStride + j] = j < nch ? t.n_N[base + fc + j] : 0;
        if (q) q[(size_t)g * kStride + j] = j < nch ? t.n_Q[base + fc + j] : 0.f;
    }
}

// ------------------------------------------------------------------ action choice (train.py:421-430)
__global__ void choose_kernel(Tree t, int32_t *actions) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= t.G) return;
    const int player = t.r_player[g];
    if ((t.g_status[g] & 1) || player == 0) { actions[g] = -1; return; }
    const size_t base = arena_base(t, g);
    const int root = t.g_root[g];
    const uint64_t me = player == 1 ? t.r_p1[g] : t.r_p2[g], opp = player == 1 ? t.r_p2[g] : t.r_p1[g];
    const uint64_t legal = legal_moves(me, opp);
    const int n = legal ? __popcll(legal) : 1;
    const int nch = t.n_meta[base + root] >> 8, fc = t.n_child[base + root];
    uint32_t r[4];
    philox(t.g_gameid[g], (uint32_t)t.g_ply[g], 0xFFFFFFFFu, 2u, t.seed_lo, t.seed_hi, r);
    int pick_j = 0;
    if (nch != 0) {
        if (t.g_ply[g] < t.temperature_plies && t.inv_temperature != 1.0f) {
            // visit_counts ** (1 / tau), normalised, then one draw (train.py:423-425; np.random.choice walks the cumulative
            // sum).  The counts are divided by the largest one first: the same distribution, and no float overflow for small
            // tau (the reference's float32 power overflows to inf there and np.random.choice raises).
            int nmax = 0;
            for (int j = 0; j < n; ++j) nmax = max(nmax, t.n_N[base + fc + j]);
            const float inv_max = nmax > 0 ? 1.0f / (float)nmax : 0.f;
            float wsum = 0.f;
            for (int j = 0; j < n; ++j) wsum += powf((float)t.n_N[base + fc + j] * inv_max, t.inv_temperature);
            if (!(wsum > 0.f)) pick_j = (int)(((uint64_t)r[0] * (uint32_t)n) >> 32);
            else {
                const float pick = ((float)(r[0] >> 8) * (1.0f / 16777216.0f)) * wsum;
                float acc = 0.f;
                pick_j = n - 1;
                for (int j = 0; j < n; ++j) { acc += powf((float)t.n_N[base + fc + j] * inv_max, t.inv_temperature); if (pick < acc) { pick_j = j; break; } }
            }
        } else if (t.g_ply[g] < t.temperature_plies) {      // tau = 1: proportional to visit counts, in integers
            uint32_t sum = 0;
            for (int j = 0; j < n; ++j) sum += (uint32_t)t.n_N[base + fc + j];
            if (sum == 0) pick_j = (int)(((uint64_t)r[0] * (uint32_t)n) >> 32);
            else {
                const uint32_t pick = (uint32_t)(((uint64_t)r[0] * sum) >> 32);
                uint32_t acc = 0;
                pick_j = n - 1;
                for (int j = 0; j < n; ++j) { acc += (uint32_t)t.n_N[base + fc + j]; if (pick < acc) { pick_j = j; break; } }
            }
        } else {                                      // uniformly random argmax
            int best = -1, ties = 0;
            for (int j = 0; j < n; ++j) { const int c = t.n_N[base + fc + j]; if (c > best) { best = c; ties = 1; } else if (c == best) ++ties; }
            uint32_t pick = (uint32_t)(((uint64_t)r[0] * (uint32_t)ties) >> 32);
            for (int j = 0; j < n; ++j) if (t.n_N[base + fc + j] == best && pick-- == 0) { pick_j = j; break; }
        }
    }
    actions[g] = legal ? nth_action(legal, pick_j) : 64;
}

__global__ void record_kernel(Tree t, const int32_t *actions, PlyRecordDev *records, int max_plies) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= t.G) return;
    const int a = actions[g], ply = t.g_ply[g];
    if (a < 0 || (t.g_status[g] & 1) || ply >= max_plies) return;
    const size_t base = arena_base(t, g);
    const int root = t.g_root[g];
    PlyRecordDev &rec = records[(size_t)g * max_plies + ply];
    rec.p1 = t.r_p1[g]; rec.p2 = t.r_p2[g]; rec.player = t.r_player[g]; rec.action = a;
    const int nch = t.n_meta[base + root] >> 8, fc = t.n_child[base + root];
    rec.n_children = nch;
    for (int j = 0; j < kStride; ++j) rec.visit_counts[j] = j < nch ? t.n_N[base + fc + j] : 0;
}

// ------------------------------------------------------------------ re-root (mcts.cpp:114-165)
// The chosen child's subtree is copied breadth-first into the game's other arena (children stay
// contiguous and ordered); everything else is dropped by switching arenas.  The old root joins
// the retired-ancestor ring that feeds the history planes.
// One BLOCK of WARPS warps per game.  A breadth-first batch is up to 32 * WARPS parents: every thread owns a parent, the
// warps' child totals are scanned through shared memory, and each warp copies the children of its 32 parents as one flat
// range with every lane busy (child k belongs to the first parent whose inclusive prefix exceeds k: a 5-step search over
// the lanes' prefixes) -- so the children land in parent order, contiguous per parent, exactly as a sequential copy would
// place them, with 32 * WARPS copies in flight.  Round 1 copied with one warp per game (0.4-0.7 ms for a single game's
// 5-20 k node subtree, 0.5 ms per ply for 1 000 games); 32 warps per game for a handful of games (the reference's
// one-game-at-a-time callers: apply_action 0.12 ms end to end), 8 warps per game otherwise (1 000 games: the rest of a ply
// 1.7 -> 0.9 ms).
constexpr int kWideApplyGames = 64;
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) apply_kernel(Tree t, const int32_t *__restrict__ actions, int32_t *unfinished) {
    __shared__ int s_wtot[32], s_woff[32], s_total;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int action = actions[g];
    if (action < 0 || (t.g_status[g] & 1)) {
        if (threadIdx.x == 0 && !(t.g_status[g] & 1)) { atomicAdd(unfinished, 1); atomicMax(unfinished + 1, t.g_count[g]); }
        return;
    }
    const int which = t.g_which[g];
    const size_t src = ((size_t)g * 2 + which) * t.cap, dst = ((size_t)g * 2 + (which ^ 1)) * t.cap;
    const int root = t.g_root[g];
    const int rm = t.n_meta[src + root];
    const int player = rm & 3, nch = rm >> 8;
    const uint64_t rp1 = t.n_p1[src + root], rp2 = t.n_p2[src + root];
    const int hn0 = t.h_n[g];
    __syncthreads();                                   // every thread has read the per-game state before thread 0 changes it
    if (threadIdx.x == 0) {                            // retire the old root into the history ring
        const size_t h = (size_t)g * kMaxHist;
        for (int k = (hn0 < kMaxHist ? hn0 : kMaxHist - 1); k > 0; --k) { t.h_p1[h + k] = t.h_p1[h + k - 1]; t.h_p2[h + k] = t.h_p2[h + k - 1]; t.h_player[h + k] = t.h_player[h + k - 1]; }
        t.h_p1[h] = rp1; t.h_p2[h] = rp2; t.h_player[h] = player;
        t.h_n[g] = hn0 < kMaxHist ? hn0 + 1 : kMaxHist;
    }
    int new_count = 1;
    int new_player;
    uint64_t np1, np2;
    if (nch == 0) {                                    // unexpanded root: fresh childless node (:140-148)
        Position s; s.player = player; s.p1 = rp1; s.p2 = rp2;
        const uint64_t me = player == 1 ? rp1 : rp2, opp = player == 1 ? rp2 : rp1;
        s.legal = legal_moves(me, opp); s.next_legal = s.legal ? 0 : legal_moves(opp, me);
        const Position c = apply_action(s, action);
        np1 = c.p1; np2 = c.p2; new_player = c.player;
        if (threadIdx.x == 0) {
            t.n_p1[dst] = np1; t.n_p2[dst] = np2; t.n_meta[dst] = new_player; t.n_parent[dst] = -1; t.n_child[dst] = -1;
            t.n_N[dst] = 0; t.n_W[dst] = 0.f; t.n_Q[dst] = 0.f; t.n_P[dst] = 1.f;
        }
    } else {
        int idx = 0;
        if (action != 0 && nch > 1) {                  // rank of the action among the legal moves (:150-157)
            const uint64_t me = player == 1 ? rp1 : rp2, opp = player == 1 ? rp2 : rp1;
            idx = __popcll(legal_moves(me, opp) & (~0ull << (64 - action)));
        }
        const size_t c = src + t.n_child[src + root] + idx;
        np1 = t.n_p1[c]; np2 = t.n_p2[c]; new_player = t.n_meta[c] & 3;
        if (threadIdx.x == 0) {
            t.n_p1[dst] = np1; t.n_p2[dst] = np2; t.n_meta[dst] = t.n_meta[c]; t.n_parent[dst] = -1;
            t.n_child[dst] = t.n_child[c];             // still an id in the OLD arena until processed below
            t.n_N[dst] = t.n_N[c]; t.n_W[dst] = t.n_W[c]; t.n_Q[dst] = t.n_Q[c]; t.n_P[dst] = t.n_P[c];
        }
        __syncthreads();
        int head = 0;
        while (head < new_count) {
            const int batch = min(WARPS * 32, new_count - head);
            const bool valid = (int)threadIdx.x < batch;
            const int my = head + (int)threadIdx.x;
            const int fc_old = valid ? t.n_child[dst + my] : -1;
            const int nc = (valid && fc_old >= 0) ? (t.n_meta[dst + my] >> 8) : 0;
            int incl = nc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 31) s_wtot[wib] = wtotal;
            __syncthreads();
            if (wib == 0) {                            // exclusive scan of the 32 warp totals
                const int v = lane < WARPS ? s_wtot[lane] : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                s_woff[lane] = inc - v;
                if (lane == 31) s_total = inc;
            }
            __syncthreads();
            const int wbase = new_count + s_woff[wib], total = s_total;
            const int off = incl - nc;
            if (nc > 0) t.n_child[dst + my] = wbase + off;
            for (int k0 = 0; k0 < wtotal; k0 += 32) {  // this warp's children as one flat range
                const int k = k0 + lane;
                int s = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const int cand = s + step;
                    const int v = __shfl_sync(0xffffffffu, incl, (cand - 1) & 31);
                    if (cand <= 31 && v <= k) s = cand;
                }
                const int f = __shfl_sync(0xffffffffu, fc_old, s), o2 = __shfl_sync(0xffffffffu, off, s);
                if (k < wtotal) {
                    const size_t a = src + f + (k - o2), d = dst + wbase + k;
                    t.n_p1[d] = t.n_p1[a]; t.n_p2[d] = t.n_p2[a]; t.n_meta[d] = t.n_meta[a]; t.n_parent[d] = head + wib * 32 + s;
                    t.n_child[d] = t.n_child[a]; t.n_N[d] = t.n_N[a]; t.n_W[d] = t.n_W[a]; t.n_Q[d] = t.n_Q[a]; t.n_P[d] = t.n_P[a];
                }
            }
            new_count += total;
            head += batch;
            __syncthreads();                           // the batch's children are the next batches' parents
        }
    }
    if (threadIdx.x == 0) {
        t.g_which[g] = which ^ 1; t.g_root[g] = 0; t.g_count[g] = new_count; t.g_ply[g] += 1;
        t.r_p1[g] = np1; t.r_p2[g] = np2; t.r_player[g] = new_player;
        if (new_player == 0) t.g_status[g] |= 1; else atomicAdd(unfinished, 1);
        atomicMax(unfinished + 1, new_count);
    }
}

// ------------------------------------------------------------------ rolling game pool
// Step 1 (one block, deterministic): finished, not yet harvested slots get a row in the store of completed
// games and -- while unplayed games remain -- the next game id, both in slot order.
__global__ void __launch_bounds__(1024) harvest_assign_kernel(Tree t, DoneStore d, int pool_total) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int done0 = d.counters[0], next0 = d.counters[1];
    for (int g0 = 0; g0 < t.G; g0 += 1024) {
        const int g = g0 + threadIdx.x;
        const int flag = (g < t.G && (t.g_status[g] & 3) == 1) ? 1 : 0;   // bit 0 finished, bit 1 harvested
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
            s_warp[lane] = w;                                     // inclusive over warps
        }
        __syncthreads();
        const int before = s_base + (warp ? s_warp[warp - 1] : 0) + incl - flag;   // finished slots before this one
        if (g < t.G) {
            int row = -1, nid = -1;
            if (flag) {
                if (done0 + before < d.capacity) row = done0 + before;
                if (next0 + before < pool_total) nid = next0 + before;
            }
            d.slot_done[g] = row; d.slot_new[g] = nid;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int fin = s_base;
        const int assigned = max(0, min(fin, pool_total - next0));
        d.counters[0] = min(done0 + fin, d.capacity);
        d.counters[1] = next0 + assigned;
        d.counters[2] = assigned;
    }
}

// Step 2 (one block per slot): copy the finished game's records into its row, then restart the slot with the
// new game (the body of reset_kernel) or leave it idle.
__global__ void __launch_bounds__(128) harvest_copy_kernel(Tree t, const PlyRecordDev *__restrict__ records, int max_plies, DoneStore d) {
    const int g = blockIdx.x;
    const int row = d.slot_done[g], nid = d.slot_new[g];
    if ((t.g_status[g] & 3) != 1) return;
    if (row >= 0) {
        const int plies = min(t.g_ply[g], max_plies);
        const uint4 *src = (const uint4 *)(records + (size_t)g * max_plies);
        uint4 *dst = (uint4 *)(d.records + (size_t)row * max_plies);
        const int n16 = plies * (int)(sizeof(PlyRecordDev) / 16);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x == 0) { d.plies[row] = plies; d.gameid[row] = t.g_gameid[g]; d.p1[row] = t.r_p1[g]; d.p2[row] = t.r_p2[g]; }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (nid < 0) { t.g_status[g] |= 2; return; }                  // pool exhausted: harvested, idle
    t.g_which[g] = 0;
    const size_t b = (size_t)g * 2 * t.cap;
    t.n_p1[b] = kInitialP1; t.n_p2[b] = kInitialP2;
    t.n_meta[b] = 1; t.n_parent[b] = -1; t.n_child[b] = -1;
    t.n_N[b] = 0; t.n_W[b] = 0.f; t.n_Q[b] = 0.f; t.n_P[b] = 1.f;
    t.g_root[g] = 0; t.g_count[g] = 1; t.g_status[g] = 0; t.g_ply[g] = 0; t.g_search[g] = 0; t.h_n[g] = 0;
    t.g_gameid[g] = t.first_game_id + (uint32_t)nid;
    t.r_p1[g] = kInitialP1; t.r_p2[g] = kInitialP2; t.r_player[g] = 1;
}

void launch_harvest(const Tree &t, const PlyRecordDev *records, int max_plies, const DoneStore &d, int pool_total, cudaStream_t s) {
    harvest_assign_kernel<<<1, 1024, 0, s>>>(t, d, pool_total); count_launch();
    harvest_copy_kernel<<<t.G, 128, 0, s>>>(t, records, max_plies, d); count_launch();
}

// ------------------------------------------------------------------ self-play sample of one root (mcts.cpp:63-112)
__global__ void self_play_data_kernel(Tree t, int g, float *features, float *policy) {
    __shared__ uint64_t s1[kMaxHist], s2[kMaxHist];
    const int lane = threadIdx.x & 31, tf = threadIdx.x >> 5;   // 8 warps = 8 symmetries
    const size_t base = arena_base(t, g);
    const int root = t.g_root[g];
    const int m = t.n_meta[base + root];
    const int player = m & 3, nch = m >> 8, fc = t.n_child[base + root];
    const int nh = min(t.H, 1 + t.h_n[g]);                      // the root, then its retired ancestors
    if (tf == 0 && lane < nh) {
        s1[lane] = lane == 0 ? t.n_p1[base + root] : t.h_p1[(size_t)g * kMaxHist + lane - 1];
        s2[lane] = lane == 0 ? t.n_p2[base + root] : t.h_p2[(size_t)g * kMaxHist + lane - 1];
    }
    __syncthreads();
    write_planes(features + (size_t)tf * (1 + 2 * t.H) * 64, t.H, nh, s1, s2, (float)player - 1.0f, tf, lane);
    float *pol = policy + tf * 65;
    for (int i = lane; i < 65; i += 32) pol[i] = 0.f;
    __syncwarp();
    int sum = 0;
    for (int j = 0; j < nch; ++j) sum += t.n_N[base + fc + j];
    if (sum == 0) sum = 1;
    const uint64_t me = player == 1 ? t.n_p1[base + root] : t.n_p2[base + root], opp = player == 1 ? t.n_p2[base + root] : t.n_p1[base + root];
    const uint64_t legal = legal_moves(me, opp);
    for (int j = lane; j < nch; j += 32) {
        const int a = legal ? nth_action(legal, j) : 64;
        pol[transform_action(a, tf)] = __fdiv_rn((float)t.n_N[base + fc + j], (float)sum);
    }
}

// ------------------------------------------------------------------ built-in exact evaluators (tests, smoke)
// One warp per board; lane l owns squares l and l + 32.  Integer weights and one correctly
// rounded division per output, so any IEEE implementation of the same formula agrees bit for bit.
__device__ __forceinline__ uint64_t plane_mask(const float *pl, int lane) {
    const uint32_t b0 = __ballot_sync(0xffffffffu, pl[lane] != 0.f), b1 = __ballot_sync(0xffffffffu, pl[lane + 32] != 0.f);
    return ((uint64_t)__brev(b0) << 32) | (uint64_t)__brev(b1);   // square i <-> bit (63 - i)
}

__device__ __forceinline__ int box_count(uint64_t occ, int sq) {
    const uint64_t s = 1ull << (63 - sq);
    const uint64_t h = s | ((s >> 1) & ~0x8080808080808080ull) | ((s << 1) & ~0x0101010101010101ull);
    return __popcll((h | (h << 8) | (h >> 8)) & occ);
}

__global__ void eval_synth_kernel(const float *__restrict__ f, int planes, size_t n, float *policy, float *value) {
    const int lane = threadIdx.x & 31;
    const size_t b = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    if (b >= n) return;
    const float *x = f + b * (size_t)planes * 64;
    const uint64_t bl = plane_mask(x + 64, lane), wh = plane_mask(x + 128, lane);
    const int w0 = 1 + box_count(bl, lane) + box_count(wh, lane), w1 = 1 + box_count(bl, lane + 32) + box_count(wh, lane + 32);
    const float sum = (float)(warp_sum(w0 + w1) + 1);
    float *po = policy + b * 65;
    po[lane] = __fdiv_rn((float)w0, sum); po[lane + 32] = __fdiv_rn((float)w1, sum);
    if (lane == 0) {
        po[64] = __fdiv_rn(1.0f, sum);
        const float side = __fsub_rn(1.0f, __fmul_rn(2.0f, x[0]));
        value[b] = __fdiv_rn(__fmul_rn(side, (float)(__popcll(bl) - __popcll(wh))), 64.0f);
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 32; z *= 0x9E3779B97F4A7C15ull; z ^= z >> 29;
    z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
    return z;
}

__global__ void eval_hash_kernel(const float *__restrict__ f, int planes, size_t n, float *policy, float *value) {
    const int lane = threadIdx.x & 31;
    const size_t b = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    if (b >= n) return;
    const float *x = f + b * (size_t)planes * 64;
    uint64_t h = 0x243F6A8885A308D3ull;
    for (int c = 0; c < planes; ++c) h = mix64(h ^ plane_mask(x + c * 64, lane));
    const uint32_t w0 = 1u + (uint32_t)(mix64(h + (uint64_t)(lane + 1) * 0xD1B54A32D192ED03ull) % 997u);
    const uint32_t w1 = 1u + (uint32_t)(mix64(h + (uint64_t)(lane + 33) * 0xD1B54A32D192ED03ull) % 997u);
    const uint32_t w64 = 1u + (uint32_t)(mix64(h + (uint64_t)65 * 0xD1B54A32D192ED03ull) % 997u);
    const float sum = (float)((uint32_t)warp_sum((int)(w0 + w1)) + w64);
    float *po = policy + b * 65;
    po[lane] = __fdiv_rn((float)w0, sum); po[lane + 32] = __fdiv_rn((float)w1, sum);
    if (lane == 0) {
        po[64] = __fdiv_rn((float)w64, sum);
        const int v = (int)(mix64(h ^ 0xA5A5A5A5A5A5A5A5ull) % 2001u) - 1000;
        value[b] = __fdiv_rn((float)v, 1000.0f);
    }
}

// ------------------------------------------------------------------ training tensors (train.py:432-450)
__global__ void __launch_bounds__(256) training_samples_kernel(const PlyRecordDev *__restrict__ records, const int32_t *__restrict__ plies,
                                                               const int64_t *__restrict__ sample_base, const uint64_t *fin_p1,
                                                               const uint64_t *fin_p2, int G, int max_plies, int H,
                                                               float *features, float *policy, float *value) {
    __shared__ uint64_t s1[kMaxHist], s2[kMaxHist];
    const int lane = threadIdx.x & 31, tf = threadIdx.x >> 5;    // block = one (game, ply), warps = symmetries
    const int g = blockIdx.x / max_plies, ply = blockIdx.x % max_plies;
    if (g >= G || ply >= plies[g] || sample_base[g] < 0) return;
    const PlyRecordDev &rec = records[(size_t)g * max_plies + ply];
    const int nh = min(H, ply + 1);
    if (threadIdx.x < nh) { s1[threadIdx.x] = records[(size_t)g * max_plies + ply - threadIdx.x].p1; s2[threadIdx.x] = records[(size_t)g * max_plies + ply - threadIdx.x].p2; }
    __syncthreads();
    const size_t row = (size_t)sample_base[g] + (size_t)ply * 8 + tf;
    write_planes(features + row * (size_t)(1 + 2 * H) * 64, H, nh, s1, s2, (float)rec.player - 1.0f, tf, lane);
    float *pol = policy + row * 65;
    for (int i = lane; i < 65; i += 32) pol[i] = 0.f;
    __syncwarp();
    int sum = 0;
    for (int j = 0; j < rec.n_children; ++j) sum += rec.visit_counts[j];
    if (sum == 0) sum = 1;
    const uint64_t me = rec.player == 1 ? rec.p1 : rec.p2, opp = rec.player == 1 ? rec.p2 : rec.p1;
    const uint64_t legal = legal_moves(me, opp);
    for (int j = lane; j < rec.n_children; j += 32) {
        const int a = legal ? nth_action(legal, j) : 64;
        pol[transform_action(a, tf)] = __fdiv_rn((float)rec.visit_counts[j], (float)sum);
    }
    if (lane == 0) {
        const int d = __popcll(fin_p1[g]) - __popcll(fin_p2[g]);
        const float z = d > 0 ? 1.f : (d < 0 ? -1.f : 0.f);      // Black's view, sign flips every ply
        value[row] = (ply & 1) ? -z : z;
    }
}

// ------------------------------------------------------------------ launchers
static inline unsigned blocks(size_t items, int per_block) { return (unsigned)((items + per_block - 1) / per_block); }
// Narrow launches are bounded to two blocks per SM: spread round-robin they never fill the registers a convolution
// CTA needs on any SM (see select_kernel).
constexpr unsigned kNarrowGrid = 2 * 148;
// An SM runs one shared-memory / L1 split at a time.  The convolution needs the largest shared-memory carve-out; a
// kernel that prefers another split cannot share the SM with it (the SM drains first), so the narrow kernels ask for
// the same carve-out although they use next to no shared memory.
static void narrow_attributes() {
    static unsigned long long done_devices = 0;
    if (!first_use_on_device(done_devices)) return;
    cudaFuncSetAttribute(select_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(encode_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(expand_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

void launch_reset(const Tree &t, cudaStream_t s) { reset_kernel<<<blocks(t.G, 128), 128, 0, s>>>(t); count_launch(); }
void launch_select(const Tree &t, cudaStream_t s, bool narrow, bool pdl) {
    if (narrow) narrow_attributes();
    if (narrow) select_kernel<1><<<std::min(blocks(t.G, 1), kNarrowGrid), 32, 0, s>>>(t);
    // Warps per game: a handful of games (the reference's one-game-at-a-time callers) have the SMs to themselves, 16 warps
    // spread a level's nodes widest (single game: 4.55 / 4.22 / 4.09 / 4.02 ms per search with 2 / 4 / 8 / 16 warps); a
    // thousand games already fill the machine, two warps per game are best there (select + encode 4.96 / 5.37 / 7.13 ms per
    // ply with 2 / 4 / 8 warps).
    else if (t.G <= 64) launch_kernel(select_kernel<16>, dim3((unsigned)t.G), dim3(512), 0, s, pdl, t);
    else if (t.G <= 512) launch_kernel(select_kernel<4>, dim3((unsigned)t.G), dim3(128), 0, s, pdl, t);
    else launch_kernel(select_kernel<2>, dim3((unsigned)t.G), dim3(64), 0, s, pdl, t);
    count_launch();
}
void launch_encode(const Tree &t, float *features, cudaStream_t s, bool narrow, void *nhwc, bool pdl) {
    if (narrow) encode_kernel<2><<<std::min(blocks((size_t)t.G * t.inflight, 2), kNarrowGrid), 64, 0, s>>>(t, features, (uint16_t *)nhwc);
    else launch_kernel(encode_kernel<4>, dim3(blocks((size_t)t.G * t.inflight, 4)), dim3(128), 0, s, pdl, t, features, (uint16_t *)nhwc);
    count_launch();
}
void launch_expand(const Tree &t, const float *policy, const float *value, cudaStream_t s, bool narrow, bool pdl) {
    if (narrow) expand_kernel<1><<<std::min(blocks(t.G, 1), kNarrowGrid), 32, 0, s>>>(t, policy, value);
    else if (t.G <= 64) launch_kernel(expand_kernel<4>, dim3((unsigned)t.G), dim3(128), 0, s, pdl, t, policy, value);
    else launch_kernel(expand_kernel<2>, dim3((unsigned)t.G), dim3(64), 0, s, pdl, t, policy, value);
    count_launch();
}
void launch_end_search(const Tree &t, cudaStream_t s) { end_search_kernel<<<blocks(t.G, 128), 128, 0, s>>>(t); count_launch(); }
void launch_root_stats(const Tree &t, int32_t *n_children, int32_t *counts, float *q, int32_t *root_n, cudaStream_t s) {
    root_stats_kernel<<<blocks((size_t)t.G * 32, 128), 128, 0, s>>>(t, n_children, counts, q, root_n); count_launch();
}
void launch_choose(const Tree &t, int32_t *actions, cudaStream_t s) { choose_kernel<<<blocks(t.G, 128), 128, 0, s>>>(t, actions); count_launch(); }
void launch_record(const Tree &t, const int32_t *actions, PlyRecordDev *records, int max_plies, cudaStream_t s) {
    record_kernel<<<blocks(t.G, 128), 128, 0, s>>>(t, actions, records, max_plies); count_launch();
}
__global__ void max_count_kernel(Tree t, int32_t *out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    int c = g < t.G ? t.g_count[g] : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, c);
}
void launch_max_count(const Tree &t, int32_t *out, cudaStream_t s) { max_count_kernel<<<blocks(t.G, 128), 128, 0, s>>>(t, out); count_launch(); }
void launch_apply(const Tree &t, const int32_t *actions, int32_t *unfinished, cudaStream_t s) {
    if (t.G <= kWideApplyGames) apply_kernel<32><<<t.G, 1024, 0, s>>>(t, actions, unfinished);
    else apply_kernel<8><<<t.G, 256, 0, s>>>(t, actions, unfinished);
    count_launch();
}
void launch_self_play_data(const Tree &t, int game, float *features, float *policy, cudaStream_t s) {
    self_play_data_kernel<<<1, 256, 0, s>>>(t, game, features, policy); count_launch();
}
void launch_eval_synth(const float *features, int planes, size_t n, float *policy, float *value, cudaStream_t s) {
    eval_synth_kernel<<<blocks(n * 32, 128), 128, 0, s>>>(features, planes, n, policy, value); count_launch();
}
void launch_eval_hash(const float *features, int planes, size_t n, float *policy, float *value, cudaStream_t s) {
    eval_hash_kernel<<<blocks(n * 32, 128), 128, 0, s>>>(features, planes, n, policy, value); count_launch();
}
void launch_training_samples(const Tree &t, const PlyRecordDev *records, const int32_t *plies, const int64_t *sample_base,
                             int max_plies, float *features, float *policy, float *value, cudaStream_t s) {
    launch_training_samples_rows(t.G, t.H, records, plies, sample_base, t.r_p1, t.r_p2, max_plies, features, policy, value, s);
}
void launch_training_samples_rows(int rows, int H, const PlyRecordDev *records, const int32_t *plies, const int64_t *sample_base,
                                  const uint64_t *fin_p1, const uint64_t *fin_p2, int max_plies, float *features, float *policy,
                                  float *value, cudaStream_t s) {
    if (rows <= 0) return;
    training_samples_kernel<<<(unsigned)((size_t)rows * max_plies), 256, 0, s>>>(records, plies, sample_base, fin_p1, fin_p2, rows,
                                                                               max_plies, H, features, policy, value);
    count_launch();
}

__global__ void copy_arenas_kernel(Tree a, Tree b) {
    const int g = blockIdx.x;
    const int which = a.g_which[g], n = a.g_count[g];
    const size_t src = ((size_t)g * 2 + which) * a.cap, dst = ((size_t)g * 2 + which) * b.cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        b.n_p1[dst + i] = a.n_p1[src + i]; b.n_p2[dst + i] = a.n_p2[src + i]; b.n_meta[dst + i] = a.n_meta[src + i];
        b.n_parent[dst + i] = a.n_parent[src + i]; b.n_child[dst + i] = a.n_child[src + i]; b.n_N[dst + i] = a.n_N[src + i];
        b.n_W[dst + i] = a.n_W[src + i]; b.n_Q[dst + i] = a.n_Q[src + i]; b.n_P[dst + i] = a.n_P[src + i];
    }
}
void launch_copy_arenas(const Tree &from, const Tree &to, cudaStream_t s) {
    copy_arenas_kernel<<<from.G, 256, 0, s>>>(from, to); count_launch();
}

}  // namespace oz
