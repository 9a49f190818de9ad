// Rules part of the C ABI: host-side Position (value type), batched rule kernels, perft.
// reference: cpp/src/include/position.h, cpp/src/include/transformation.h.
#include <bitset>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "oz_common.cuh"
#include "oz_rules.cuh"

namespace oz {

static thread_local std::string t_error;
unsigned long long g_kernel_launches = 0;

void set_error(const std::string &msg) { t_error = msg; }
int fail(int status, const std::string &msg) { t_error = msg; return status; }

static inline Position to_pos(const oz_position *p) {
    Position q; q.player = p->player; q.p1 = p->p1; q.p2 = p->p2; q.legal = p->legal; q.next_legal = p->next_legal;
    return q;
}
static inline void from_pos(const Position &q, oz_position *p) {
    p->player = q.player; p->p1 = q.p1; p->p2 = q.p2; p->legal = q.legal; p->next_legal = q.next_legal;
}

// ------------------------------------------------------------------ batch kernels

__global__ void movegen_kernel(const uint64_t *__restrict__ player, const uint64_t *__restrict__ opp,
                               uint64_t *__restrict__ legal, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        legal[i] = legal_moves(player[i], opp[i]);
}

__global__ void flips_kernel(const uint64_t *__restrict__ move, const uint64_t *__restrict__ player,
                             const uint64_t *__restrict__ opp, uint64_t *__restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = flips(move[i], player[i], opp[i]);
}

// Positions travel as structure-of-arrays on the device so every access is coalesced.
__global__ void apply_action_kernel(const int32_t *__restrict__ player, const uint64_t *__restrict__ p1,
                                    const uint64_t *__restrict__ p2, const uint64_t *__restrict__ next_legal,
                                    const int32_t *__restrict__ action, int32_t *__restrict__ o_player,
                                    uint64_t *__restrict__ o_p1, uint64_t *__restrict__ o_p2,
                                    uint64_t *__restrict__ o_legal, uint64_t *__restrict__ o_next, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        Position s; s.player = player[i]; s.p1 = p1[i]; s.p2 = p2[i]; s.legal = 0; s.next_legal = next_legal[i];
        Position r = apply_action(s, action[i]);
        o_player[i] = r.player; o_p1[i] = r.p1; o_p2[i] = r.p2; o_legal[i] = r.legal; o_next[i] = r.next_legal;
    }
}

// One warp per item: lane l writes squares l and l+32 of every plane (transformation.h:83-116).
__global__ void features_kernel(const uint64_t *__restrict__ h_p1, const uint64_t *__restrict__ h_p2,
                                const int32_t *__restrict__ cur_player, const int32_t *__restrict__ n_hist,
                                const int32_t *__restrict__ tf, int H, float *__restrict__ out, size_t n) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    if (warp >= n) return;
    const int t = tf[warp];
    float *f = out + warp * (size_t)(1 + 2 * H) * 64;
    const float side = (float)cur_player[warp] - 1.0f;
    f[lane] = side; f[lane + 32] = side;
    // Destination square j reads source square i with transform(i) == j.  Writing by source
    // (scatter) keeps the loop identical to the reference; both squares of a lane are distinct.
    const int nh = n_hist[warp];
    for (int k = 0; k < H; ++k) {
        float *pl = f + 64 + k * 128;
        if (k >= nh) {
            pl[lane] = 0.f; pl[lane + 32] = 0.f; pl[64 + lane] = 0.f; pl[96 + lane] = 0.f;
        } else {
            const uint64_t b = h_p1[warp * H + k], w = h_p2[warp * H + k];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int i = lane + 32 * s;
                const int j = transform_action(i, t);
                const uint64_t m = 1ull << (63 - i);
                pl[j] = (b & m) ? 1.f : 0.f;
                pl[64 + j] = (w & m) ? 1.f : 0.f;
            }
        }
    }
}

// ------------------------------------------------------------------ perft

struct FNode { uint64_t me, opp; };  // side to move first

__device__ __forceinline__ void warp_add(unsigned long long *dst, unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, v);
}

// Children per frontier node (0 for a terminal node, which is a leaf at any depth).
__device__ __forceinline__ int child_count(const FNode &nd, uint64_t &legal) {
    legal = legal_moves(nd.me, nd.opp);
    if (legal) return __popcll(legal);
    return legal_moves(nd.opp, nd.me) ? 1 : 0;
}

// counters[0] = children of this level, counters[1] = terminal leaves met
__global__ void bfs_count_kernel(const FNode *__restrict__ in, size_t n, unsigned long long *counters) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned long long c = 0, t = 0;
    if (i < n) { uint64_t legal; int k = child_count(in[i], legal); c = k; t = (k == 0); }
    warp_add(&counters[0], c);
    warp_add(&counters[1], t);
}

__global__ void bfs_expand_kernel(const FNode *__restrict__ in, size_t n, FNode *__restrict__ out,
                                  unsigned long long *cursor) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    FNode nd = {0, 0};
    uint64_t legal = 0;
    int k = 0;
    if (i < n) { nd = in[i]; k = child_count(nd, legal); }
    // Warp-aggregated reservation: one atomic per warp.
    int incl = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && total) base = atomicAdd(cursor, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 0) + (unsigned long long)(incl - k);
    if (k == 0) return;
    if (legal == 0) { out[base] = FNode{nd.opp, nd.me}; return; }  // pass
    while (legal) {
        const uint64_t mv = legal & (~legal + 1);
        legal ^= mv;
        const uint64_t f = flips(mv, nd.me, nd.opp);
        out[base++] = FNode{nd.opp & ~f, nd.me | mv | f};
    }
}

// Depth-first walk of R more plies below a fully formed position.  nodes counts the
// positions R plies down (terminal positions met earlier count once); mob adds their
// child counts, i.e. the leaf count one ply further.
template <int R>
__device__ __forceinline__ void walk(uint64_t me, uint64_t opp, unsigned long long &nodes, unsigned long long &mob) {
    const uint64_t legal = legal_moves(me, opp);
    if constexpr (R == 0) {
        nodes += 1;
        mob += legal ? (unsigned)__popcll(legal) : 1u;
    } else {
        if (legal == 0) {
            if (legal_moves(opp, me) == 0) { nodes += 1; mob += 1; return; }
            walk<R - 1>(opp, me, nodes, mob);
            return;
        }
        uint64_t rest = legal;
        while (rest) {
            const uint64_t mv = rest & (~rest + 1);
            rest ^= mv;
            const uint64_t f = flips(mv, me, opp);
            walk<R - 1>(opp & ~f, me | mv | f, nodes, mob);
        }
    }
}

// Issue-rate micro-benchmark of the integer pipe the move generator lives on: eight independent chains of LOP3 (three-input
// logic) and SHF (funnel shift), the two instructions a 64-bit Kogge-Stone fill compiles to, with nothing else in the
// loop.  thread-operations per second of this kernel = the roofline denominator of perft.
constexpr int kIntPeakOpsPerIter = 32;
#define OZ_LOP3(d, a, b, c) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define OZ_SHF(d, a, b) asm volatile("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(d) : "r"(a), "r"(b))
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, int iters) {
    uint32_t a0 = threadIdx.x, a1 = blockIdx.x, a2 = 0x9E3779B9u, a3 = 0x7F4A7C15u, a4 = a0 * 3u, a5 = a1 * 5u, a6 = ~a0, a7 = ~a1;
    const uint32_t k = 0xBF58476Du ^ (uint32_t)iters;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {            // 2 x (8 LOP3 + 8 SHF) = 32 operations, eight independent chains
            OZ_LOP3(a0, a0, a4, k); OZ_LOP3(a1, a1, a5, k); OZ_LOP3(a2, a2, a6, k); OZ_LOP3(a3, a3, a7, k);
            OZ_LOP3(a4, a4, a0, k); OZ_LOP3(a5, a5, a1, k); OZ_LOP3(a6, a6, a2, k); OZ_LOP3(a7, a7, a3, k);
            OZ_SHF(a0, a0, a1); OZ_SHF(a1, a1, a2); OZ_SHF(a2, a2, a3); OZ_SHF(a3, a3, a0);
            OZ_SHF(a4, a4, a5); OZ_SHF(a5, a5, a6); OZ_SHF(a6, a6, a7); OZ_SHF(a7, a7, a4);
        }
    }
    const uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x1234567u) *out = r;               // keeps the chains alive; practically never true
}
#undef OZ_LOP3
#undef OZ_SHF

template <int R>
__global__ void __launch_bounds__(128) dfs_kernel(const FNode *__restrict__ in, size_t n, unsigned long long *counters) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned long long nodes = 0, mob = 0;
    if (i < n) { const FNode nd = in[i]; walk<R>(nd.me, nd.opp, nodes, mob); }
    warp_add(&counters[2], nodes);
    warp_add(&counters[3], mob);
}

static int launch_dfs(int R, const FNode *in, size_t n, unsigned long long *counters) {
    const unsigned grid = (unsigned)((n + 127) / 128);
    switch (R) {
        case 0: dfs_kernel<0><<<grid, 128>>>(in, n, counters); break;
        case 1: dfs_kernel<1><<<grid, 128>>>(in, n, counters); break;
        case 2: dfs_kernel<2><<<grid, 128>>>(in, n, counters); break;
        case 3: dfs_kernel<3><<<grid, 128>>>(in, n, counters); break;
        case 4: dfs_kernel<4><<<grid, 128>>>(in, n, counters); break;
        default: return fail(OZ_ERR_INVALID_ARGUMENT, "internal: dfs depth");
    }
    count_launch();
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

}  // namespace oz

using namespace oz;

extern "C" {

const char *oz_last_error(void) { return t_error.c_str(); }
int oz_abi_version(void) { return 4; }

uint64_t oz_get_legal_moves(uint64_t a, uint64_t b) { return legal_moves(a, b); }
uint64_t oz_get_flips(uint64_t m, uint64_t a, uint64_t b) { return flips(m, a, b); }
int oz_transform_action(int a, int t) { return transform_action(a, t); }

int oz_position_initial(oz_position *out) { from_pos(initial_position(), out); return OZ_OK; }

int oz_position_at(const oz_position *p, int index, int *square) {
    if (!(0 <= index && index < 64))  // position.h:285-292
        return fail(OZ_ERR_OUT_OF_RANGE, "Expected 0 <= index < 64, but got " + std::to_string(index) + ".");
    const uint64_t m = 1ull << (63 - index);
    *square = (p->p1 & m) ? 1 : ((p->p2 & m) ? 2 : 0);
    return OZ_OK;
}

int oz_position_is_legal_move(const oz_position *p, int index, int *legal) {
    if (!(0 <= index && index < 64))  // position.h:299-306
        return fail(OZ_ERR_OUT_OF_RANGE, "Expected 0 <= index < 64, but got " + std::to_string(index) + ".");
    *legal = (p->legal & (1ull << (63 - index))) != 0;
    return OZ_OK;
}

int oz_position_legal_actions(const oz_position *p, int *actions, int *count) {
    int n = 0;  // position.h:308-326
    if (p->player != 0) {
        if (p->legal == 0) actions[n++] = 64;
        else for (int a = 0; a < 64; ++a) if (p->legal & (1ull << (63 - a))) actions[n++] = a;
    }
    *count = n;
    return OZ_OK;
}

int oz_position_apply_move(const oz_position *p, uint64_t mv, oz_position *out) {
    if (__builtin_popcountll(mv) != 1)  // position.h:367-372
        return fail(OZ_ERR_INVALID_ARGUMENT,
                    "Expected a single bit in move_mask, but got 0b" + std::bitset<64>(mv).to_string() + ".");
    if ((mv & p->legal) == 0)           // position.h:373-378
        return fail(OZ_ERR_INVALID_ARGUMENT, "0b" + std::bitset<64>(mv).to_string() + " is not a legal move.");
    from_pos(apply_move(to_pos(p), mv), out);
    return OZ_OK;
}

int oz_position_apply_pass(const oz_position *p, oz_position *out) {
    if (p->player == 0)  // position.h:389-393
        return fail(OZ_ERR_INVALID_ARGUMENT, "Pass is not allowed in a terminal position.");
    if (p->legal != 0)   // position.h:394-398
        return fail(OZ_ERR_INVALID_ARGUMENT, "Pass is not allowed when there are legal moves.");
    from_pos(apply_pass(to_pos(p)), out);
    return OZ_OK;
}

int oz_position_apply_action(const oz_position *p, int action, oz_position *out) {
    if (!(0 <= action && action < 65))  // position.h:412-416
        return fail(OZ_ERR_OUT_OF_RANGE, "Expected 0 <= action < 65, but got " + std::to_string(action) + ".");
    if (action == 64) return oz_position_apply_pass(p, out);
    const uint64_t mv = 1ull << (63 - action);
    if ((mv & p->legal) == 0)           // position.h:421-425
        return fail(OZ_ERR_INVALID_ARGUMENT, std::to_string(action) + " is not a legal action.");
    from_pos(apply_move(to_pos(p), mv), out);
    return OZ_OK;
}

int oz_position_to_string(const oz_position *p, char *buf, size_t buf_size) {
    std::string s = "  a b c d e f g h\n";  // position.h:429-456
    uint64_t m = 1ull << 63;
    for (int row = 0; row < 8; ++row) {
        s.push_back((char)('1' + row));
        for (int col = 0; col < 8; ++col) {
            s.push_back(' ');
            if (m & p->p1) s += "●";
            else if (m & p->p2) s += "○";
            else if (m & p->legal) s += "×";
            else s += "·";
            m >>= 1;
        }
        if (row < 7) s.push_back('\n');
    }
    if (s.size() + 1 > buf_size) return fail(OZ_ERR_INVALID_ARGUMENT, "buffer too small");
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return OZ_OK;
}

// ------------------------------------------------------------------ batched rules

int oz_movegen_batch_device(const uint64_t *player, const uint64_t *opp, uint64_t *legal, size_t n, void *stream) {
    if (n == 0) return OZ_OK;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
    movegen_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(player, opp, legal, n);
    count_launch();
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

int oz_movegen_batch(const uint64_t *player, const uint64_t *opp, uint64_t *legal, size_t n) {
    if (n == 0) return OZ_OK;
    uint64_t *d = nullptr;
    OZ_CUDA(cudaMalloc(&d, 3 * n * sizeof(uint64_t)));
    int st = OZ_OK;
    do {
        if (cudaMemcpy(d, player, n * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d + n, opp, n * 8, cudaMemcpyHostToDevice) != cudaSuccess) { st = fail(OZ_ERR_CUDA, "H2D copy failed"); break; }
        st = oz_movegen_batch_device(d, d + n, d + 2 * n, n, nullptr);
        if (st) break;
        cudaError_t e = cudaMemcpy(legal, d + 2 * n, n * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) st = fail(OZ_ERR_CUDA, std::string("movegen: ") + cudaGetErrorString(e));
    } while (0);
    cudaFree(d);
    return st;
}

int oz_flips_batch(const uint64_t *move, const uint64_t *player, const uint64_t *opp, uint64_t *out, size_t n) {
    if (n == 0) return OZ_OK;
    uint64_t *d = nullptr;
    OZ_CUDA(cudaMalloc(&d, 4 * n * sizeof(uint64_t)));
    int st = OZ_OK;
    do {
        if (cudaMemcpy(d, move, n * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d + n, player, n * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d + 2 * n, opp, n * 8, cudaMemcpyHostToDevice) != cudaSuccess) { st = fail(OZ_ERR_CUDA, "H2D copy failed"); break; }
        const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
        flips_kernel<<<grid, 256>>>(d, d + n, d + 2 * n, d + 3 * n, n);
        count_launch();
        cudaError_t e = cudaMemcpy(out, d + 3 * n, n * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) st = fail(OZ_ERR_CUDA, std::string("flips: ") + cudaGetErrorString(e));
    } while (0);
    cudaFree(d);
    return st;
}

int oz_apply_action_batch(const oz_position *in, const int32_t *actions, oz_position *out, size_t n) {
    if (n == 0) return OZ_OK;
    // AoS (ABI) <-> SoA (device) on the host side of the copy.
    std::vector<int32_t> pl(n), act(actions, actions + n);
    std::vector<uint64_t> p1(n), p2(n), nl(n);
    for (size_t i = 0; i < n; ++i) { pl[i] = in[i].player; p1[i] = in[i].p1; p2[i] = in[i].p2; nl[i] = in[i].next_legal; }
    char *d = nullptr;
    const size_t b8 = n * 8, b4 = n * 4;
    OZ_CUDA(cudaMalloc(&d, 7 * b8 + 3 * b4));
    uint64_t *d_p1 = (uint64_t *)d, *d_p2 = d_p1 + n, *d_nl = d_p2 + n, *o_p1 = d_nl + n, *o_p2 = o_p1 + n,
             *o_lg = o_p2 + n, *o_nl = o_lg + n;
    int32_t *d_pl = (int32_t *)(o_nl + n), *d_act = d_pl + n, *o_pl = d_act + n;
    int st = OZ_OK;
    do {
        if (cudaMemcpy(d_p1, p1.data(), b8, cudaMemcpyHostToDevice) || cudaMemcpy(d_p2, p2.data(), b8, cudaMemcpyHostToDevice) ||
            cudaMemcpy(d_nl, nl.data(), b8, cudaMemcpyHostToDevice) || cudaMemcpy(d_pl, pl.data(), b4, cudaMemcpyHostToDevice) ||
            cudaMemcpy(d_act, act.data(), b4, cudaMemcpyHostToDevice)) { st = fail(OZ_ERR_CUDA, "H2D copy failed"); break; }
        const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
        apply_action_kernel<<<grid, 256>>>(d_pl, d_p1, d_p2, d_nl, d_act, o_pl, o_p1, o_p2, o_lg, o_nl, n);
        count_launch();
        std::vector<uint64_t> r1(n), r2(n), rl(n), rn(n);
        if (cudaMemcpy(r1.data(), o_p1, b8, cudaMemcpyDeviceToHost) || cudaMemcpy(r2.data(), o_p2, b8, cudaMemcpyDeviceToHost) ||
            cudaMemcpy(rl.data(), o_lg, b8, cudaMemcpyDeviceToHost) || cudaMemcpy(rn.data(), o_nl, b8, cudaMemcpyDeviceToHost) ||
            cudaMemcpy(pl.data(), o_pl, b4, cudaMemcpyDeviceToHost)) { st = fail(OZ_ERR_CUDA, std::string("apply: ") + cudaGetErrorString(cudaGetLastError())); break; }
        for (size_t i = 0; i < n; ++i) { out[i].player = pl[i]; out[i].p1 = r1[i]; out[i].p2 = r2[i]; out[i].legal = rl[i]; out[i].next_legal = rn[i]; }
    } while (0);
    cudaFree(d);
    return st;
}

int oz_features_batch(const oz_position *history, const int32_t *n_hist, const int32_t *tf, int H,
                      float *features, size_t n) {
    if (n == 0) return OZ_OK;
    if (H < 1) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected history_size >= 1, but got " + std::to_string(H) + ".");
    std::vector<uint64_t> p1(n * H), p2(n * H);
    std::vector<int32_t> cur(n);
    for (size_t i = 0; i < n; ++i) {
        cur[i] = history[i * H].player;
        for (int k = 0; k < H; ++k) { p1[i * H + k] = history[i * H + k].p1; p2[i * H + k] = history[i * H + k].p2; }
    }
    const size_t fb = n * (size_t)(1 + 2 * H) * 64 * sizeof(float);
    char *d = nullptr;
    OZ_CUDA(cudaMalloc(&d, 2 * n * H * 8 + 3 * n * 4 + fb));
    uint64_t *d_p1 = (uint64_t *)d, *d_p2 = d_p1 + n * H;
    float *d_f = (float *)(d_p2 + n * H);
    int32_t *d_cur = (int32_t *)((char *)d_f + fb), *d_nh = d_cur + n, *d_tf = d_nh + n;
    int st = OZ_OK;
    do {
        if (cudaMemcpy(d_p1, p1.data(), n * H * 8, cudaMemcpyHostToDevice) || cudaMemcpy(d_p2, p2.data(), n * H * 8, cudaMemcpyHostToDevice) ||
            cudaMemcpy(d_cur, cur.data(), n * 4, cudaMemcpyHostToDevice) || cudaMemcpy(d_nh, n_hist, n * 4, cudaMemcpyHostToDevice) ||
            cudaMemcpy(d_tf, tf, n * 4, cudaMemcpyHostToDevice)) { st = fail(OZ_ERR_CUDA, "H2D copy failed"); break; }
        const unsigned grid = (unsigned)((n * 32 + 255) / 256);
        features_kernel<<<grid, 256>>>(d_p1, d_p2, d_cur, d_nh, d_tf, H, d_f, n);
        count_launch();
        cudaError_t e = cudaMemcpy(features, d_f, fb, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) st = fail(OZ_ERR_CUDA, std::string("features: ") + cudaGetErrorString(e));
    } while (0);
    cudaFree(d);
    return st;
}

// ------------------------------------------------------------------ integer-pipe issue peak (roofline denominator of perft)

int oz_debug_int_peak(double *ops_per_second, double *device_ms) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t *d_out = nullptr;
    OZ_CUDA(cudaMalloc(&d_out, sizeof(uint32_t)));
    const int iters = 4096, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {            // first launch warms up; best of the rest
        cudaEventRecord(e0);
        int_peak_kernel<<<blocks, threads>>>(d_out, iters); count_launch();
        cudaEventRecord(e1);
        cudaError_t err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) { cudaFree(d_out); return fail(OZ_ERR_CUDA, std::string("int_peak: ") + cudaGetErrorString(err)); }
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    const double ops = (double)blocks * threads * iters * kIntPeakOpsPerIter;
    if (ops_per_second) *ops_per_second = ops / (best * 1e-3);
    if (device_ms) *device_ms = best;
    return OZ_OK;
}

// ------------------------------------------------------------------ perft

int oz_perft(const oz_position *root, int depth, int mode, uint64_t *nodes, double *device_ms) {
    if (depth < 0) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected depth >= 0, but got " + std::to_string(depth) + ".");
    if (mode != 0 && mode != 1) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected mode 0 or 1.");
    if (device_ms) *device_ms = 0.0;
    if (root->player == 0 || depth == 0) { *nodes = 1; return OZ_OK; }
    // mode 1 walks one ply less and returns the child-count sum of that level.
    const int walk_depth = mode == 1 ? depth - 1 : depth;
    const int R = walk_depth >= 12 ? 4 : 3;            // plies each thread walks depth-first
    const int bfs_levels = walk_depth > R ? walk_depth - R : 0;
    const int dfs_R = walk_depth - bfs_levels;

    FNode h_root = (root->player == 1) ? FNode{root->p1, root->p2} : FNode{root->p2, root->p1};
    unsigned long long *d_cnt = nullptr;   // [0] children, [1] terminals, [2] nodes, [3] mob, [4] cursor
    FNode *cur = nullptr, *next = nullptr;
    size_t cap_cur = 0, cap_next = 0;      // the two frontier buffers are kept across levels and only grow (no malloc / free per level)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int st = OZ_OK;
    unsigned long long early_leaves = 0;   // terminal positions above the walk depth
    size_t n = 1;
#define PF_CUDA(x) { cudaError_t _e = (x); if (_e != cudaSuccess) { st = fail(OZ_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); break; } }
    do {
        PF_CUDA(cudaMalloc(&d_cnt, 8 * sizeof(unsigned long long)));
        cap_cur = cap_next = (size_t)1 << 16;
        PF_CUDA(cudaMalloc(&cur, cap_cur * sizeof(FNode)));
        PF_CUDA(cudaMalloc(&next, cap_next * sizeof(FNode)));
        PF_CUDA(cudaMemcpy(cur, &h_root, sizeof(FNode), cudaMemcpyHostToDevice));
        PF_CUDA(cudaEventCreate(&ev0)); PF_CUDA(cudaEventCreate(&ev1));
        PF_CUDA(cudaEventRecord(ev0));
        for (int lvl = 0; lvl < bfs_levels && n > 0; ++lvl) {
            unsigned long long h[2];
            PF_CUDA(cudaMemsetAsync(d_cnt, 0, 8 * sizeof(unsigned long long)));
            const unsigned grid = (unsigned)((n + 255) / 256);
            bfs_count_kernel<<<grid, 256>>>(cur, n, d_cnt); count_launch();
            PF_CUDA(cudaMemcpy(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost));
            early_leaves += h[1];
            if (h[0] > cap_next) {
                cudaFree(next); next = nullptr;
                cap_next = std::max<size_t>(h[0], 4 * cap_next);
                PF_CUDA(cudaMalloc(&next, cap_next * sizeof(FNode)));
            }
            bfs_expand_kernel<<<grid, 256>>>(cur, n, next, d_cnt + 4); count_launch();
            PF_CUDA(cudaGetLastError());
            std::swap(cur, next); std::swap(cap_cur, cap_next); n = h[0];      // stream order keeps the levels apart: no synchronisation
        }
        if (st) break;
        unsigned long long h[4] = {0, 0, 0, 0};
        PF_CUDA(cudaMemsetAsync(d_cnt, 0, 8 * sizeof(unsigned long long)));
        if (n > 0) { st = launch_dfs(dfs_R, cur, n, d_cnt); if (st) break; }
        PF_CUDA(cudaEventRecord(ev1));
        PF_CUDA(cudaMemcpy(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost));
        float ms = 0.f;
        PF_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        if (device_ms) *device_ms = ms;
        *nodes = (mode == 1 ? h[3] : h[2]) + early_leaves;
    } while (0);
#undef PF_CUDA
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    cudaFree(cur); cudaFree(next); cudaFree(d_cnt);
    return st;
}

}  // extern "C"
