// Engine part of the C ABI: the MCTS facade of the reference (cpp/src/include/mcts.h,
// cpp/src/lib/mcts.cpp) for `num_games` lockstep games whose trees live in HBM.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "oz_common.cuh"
#include "oz_resnet.cuh"
#include "oz_rules.cuh"
#include "oz_tree.cuh"

using namespace oz;

static_assert(sizeof(oz_ply_record) == sizeof(PlyRecordDev), "record layout");
static_assert(OZ_MAX_CHILDREN == kStride, "stride");

struct oz_resnet {
    int device = 0;
    ResNet net;
    float *d_f = nullptr, *d_p = nullptr, *d_v = nullptr;
    size_t cap = 0;
};

constexpr int kMaxGroups = 8;             // game groups of the pipelined search / branches of a graph-replayed search

struct oz_engine {
    oz_config cfg{};
    Tree t{};
    cudaStream_t stream = nullptr;
    int inflight_cap = 0, planes_cap = 0;
    float *d_features = nullptr, *d_policy = nullptr, *d_value = nullptr;
    float *d_rate = nullptr, *d_gamma = nullptr;
    int32_t *d_actions = nullptr, *d_unfinished = nullptr;
    int32_t *d_stat_n = nullptr, *d_stat_counts = nullptr, *d_stat_rootn = nullptr;
    float *d_stat_q = nullptr;
    PlyRecordDev *d_records = nullptr;
    int max_plies = 0;
    float *d_spd_f = nullptr, *d_spd_p = nullptr;
    DoneStore done{};                     // rolling game pool (oz_engine_set_game_pool); capacity 0 = off
    int pool_total = 0;
    ResNet net;
    bool in_search = false;
    int round = 0;
    cudaGraphExec_t round_graph = nullptr;   // every round of a search (select, encode, network, expand) for launch-bound batches
    int graph_rounds = 0;                    // rounds captured in round_graph
    int graph_branches = 1;                  // game groups captured as parallel branches of round_graph
    unsigned long long graph_launches = 0;   // kernels inside round_graph
    bool warmed = false;
    int graph_mode = 1;                      // 0 never, 1 when the leaf batch is launch-bound, 2 always                     // an eager search has run with the current allocations
    // Pipelined search: the games are split into groups that run select -> network -> expand on their own streams, so
    // one group's tree kernels run underneath another group's convolutions (results do not depend on the grouping).
    int pipeline = 0;                     // 0 auto, 1 off, 2..kMaxGroups groups
    bool pipeline_pdl = true;             // keep programmatic dependent launch between conv layers while pipelined
    cudaStream_t gstream[kMaxGroups] = {};     // tree streams (low priority), one per group
    cudaStream_t nstream = nullptr;            // network stream (high priority), shared: the conv layers of all groups stay one chain
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxGroups] = {}, ev_sel[kMaxGroups] = {}, ev_net[kMaxGroups] = {};
    int active = 0;                       // games not finished
    long tree_bound = 1;                  // upper bound on the nodes of the largest live tree
    long rootn_bound = 0;                 // upper bound on any node's visit count (index into the exploration-rate table)
    oz_counters ctr{};
    // Pinned host mirrors of the small per-ply read-backs (error bits, unfinished games, largest tree, games assigned by
    // the harvest, chosen actions, root visit counts): the copies into them are truly asynchronous, so a ply blocks the
    // host exactly once, at its end (a cudaMemcpyAsync into pageable memory is a host-blocking call of its own).
    bool defer_error_check = false;       // inside oz_engine_play_ply
    size_t prof_events = 0;               // profiling events of the last search still to be read (after a synchronisation)
    bool prof_pipelined = false, prof_net = false;
    int32_t *h_pin = nullptr;             // [8]: 0 error bits, 1 unfinished, 2 largest tree, 3 assigned
    int32_t *h_actions = nullptr, *h_counts = nullptr;   // [G], [G * kStride]
    bool profiling = false;
    std::vector<cudaEvent_t> ev;          // phase boundaries of one search
};

namespace {

// ---- validation, same texts as mcts.cpp:167-241 ----
int check_ge1(const char *name, int v) {
    if (v < 1) return fail(OZ_ERR_INVALID_ARGUMENT, std::string("Expected ") + name + " >= 1, but got " + std::to_string(v) + ".");
    return OZ_OK;
}
int check_config(const oz_config &c) {
    OZ_TRY(check_ge1("history_size", c.history_size));
    OZ_TRY(check_ge1("num_simulations", c.num_simulations));
    OZ_TRY(check_ge1("num_threads", c.num_threads));
    OZ_TRY(check_ge1("batch_size", c.batch_size));
    if (!(c.c_puct_base > 0.0f)) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected c_puct_base > 0.0, but got " + std::to_string(c.c_puct_base) + ".");
    if (!(c.c_puct_init >= 0.0f)) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected c_puct_init >= 0.0, but got " + std::to_string(c.c_puct_init) + ".");
    if (!(0.0f <= c.dirichlet_epsilon && c.dirichlet_epsilon <= 1.0f))
        return fail(OZ_ERR_INVALID_ARGUMENT, "Expected 0.0 <= dirichlet_epsilon <= 1.0, but got " + std::to_string(c.dirichlet_epsilon) + ".");
    if (!(c.dirichlet_alpha >= 0.0f)) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected dirichlet_alpha >= 0.0, but got " + std::to_string(c.dirichlet_alpha) + ".");
    return OZ_OK;
}
int check_engine_limits(const oz_config &c) {
    if (c.num_games < 1) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected num_games >= 1, but got " + std::to_string(c.num_games) + ".");
    if (c.history_size > kMaxHist) return fail(OZ_ERR_INVALID_ARGUMENT, "history_size above " + std::to_string(kMaxHist) + " is not supported by the device engine.");
    if (!(c.self_play_temperature > 0.0f)) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected self_play_temperature > 0.0, but got " + std::to_string(c.self_play_temperature) + ".");
    return OZ_OK;
}

template <typename T>
int dalloc(T **p, size_t n) {
    *p = nullptr;
    OZ_CUDA(cudaMalloc((void **)p, std::max<size_t>(1, n) * sizeof(T)));
    return OZ_OK;
}

int rounds_of(const oz_config &c) {
    const int inflight = c.num_threads * c.batch_size;
    return (c.num_simulations + inflight - 1) / inflight;   // search_thread.cpp:48-50
}

constexpr int kMaxChildrenPerNode = 33;   // most legal moves an Othello position can have

// Nodes one search can add to a tree: every simulation expands at most one leaf.
long search_growth(const oz_config &c) {
    return (long)rounds_of(c) * c.num_threads * c.batch_size * kMaxChildrenPerNode;
}

// Initial arena size: room for a typical carried-over tree plus one worst-case search.  Arenas grow on demand
// (ensure_capacity), so this is a starting point, not a limit.
int default_cap(const oz_config &c) {
    if (c.max_nodes_per_game > 0) return c.max_nodes_per_game;
    const long sims = (long)rounds_of(c) * c.num_threads * c.batch_size;
    return (int)std::min<long>(48 * std::max<long>(sims, 64) + search_growth(c) + 1024, 1 << 22);
}

int alloc_pools(Tree &t, int G, int cap) {
    const size_t n = (size_t)G * 2 * cap;
    OZ_TRY(dalloc(&t.n_p1, n)); OZ_TRY(dalloc(&t.n_p2, n));
    OZ_TRY(dalloc(&t.n_meta, n)); OZ_TRY(dalloc(&t.n_parent, n)); OZ_TRY(dalloc(&t.n_child, n)); OZ_TRY(dalloc(&t.n_N, n));
    OZ_TRY(dalloc(&t.n_W, n)); OZ_TRY(dalloc(&t.n_Q, n)); OZ_TRY(dalloc(&t.n_P, n));
    t.cap = cap;
    return OZ_OK;
}
void free_pools(Tree &t) {
    cudaFree(t.n_p1); cudaFree(t.n_p2); cudaFree(t.n_meta); cudaFree(t.n_parent); cudaFree(t.n_child); cudaFree(t.n_N);
    cudaFree(t.n_W); cudaFree(t.n_Q); cudaFree(t.n_P);
    t.n_p1 = t.n_p2 = nullptr; t.n_meta = t.n_parent = t.n_child = t.n_N = nullptr; t.n_W = t.n_Q = t.n_P = nullptr;
}

void drop_graph(oz_engine *e);

int alloc_slots(oz_engine *e) {
    drop_graph(e);
    Tree &t = e->t;
    const int inflight = e->cfg.num_threads * e->cfg.batch_size, planes = 1 + 2 * e->cfg.history_size;
    if (inflight > e->inflight_cap) {
        cudaFree(t.s_leaf); cudaFree(t.s_tf); cudaFree(t.s_len); cudaFree(t.s_path); cudaFree(e->d_policy); cudaFree(e->d_value);
        const size_t slots = (size_t)t.G * inflight;
        OZ_TRY(dalloc(&t.s_leaf, slots)); OZ_TRY(dalloc(&t.s_tf, slots)); OZ_TRY(dalloc(&t.s_len, slots));
        OZ_TRY(dalloc(&t.s_path, slots * kMaxDepth));
        OZ_TRY(dalloc(&e->d_policy, slots * 65)); OZ_TRY(dalloc(&e->d_value, slots));
        OZ_CUDA(cudaMemset(t.s_tf, 0, slots * sizeof(int32_t)));
        OZ_CUDA(cudaMemset(e->d_policy, 0, slots * 65 * sizeof(float)));
        OZ_CUDA(cudaMemset(e->d_value, 0, slots * sizeof(float)));
    }
    if (inflight > e->inflight_cap || planes > e->planes_cap) {
        cudaFree(e->d_features);
        e->inflight_cap = std::max(e->inflight_cap, inflight);
        e->planes_cap = std::max(e->planes_cap, planes);
        // padded: the network evaluates whole 32-board pairs of units
        OZ_TRY(dalloc(&e->d_features, ((size_t)t.G * e->inflight_cap + kBoardPad) * e->planes_cap * 64));
        cudaFree(e->d_spd_f);
        OZ_TRY(dalloc(&e->d_spd_f, (size_t)8 * e->planes_cap * 64));
    }
    t.inflight = inflight;
    t.H = e->cfg.history_size;
    return OZ_OK;
}

// Exploration rate by visit count, computed with the host libm the reference links
// (search_thread.cpp:198-203): logf((1 + N + c_base) / c_base) + c_init.
int build_rate_table(oz_engine *e, long min_len = 0) {
    const long sims = (long)rounds_of(e->cfg) * e->cfg.num_threads * e->cfg.batch_size;
    const long len = std::min<long>(std::max<long>({sims * 140 + 1024, 4096L, min_len, (long)e->t.rate_len}), 1L << 26);   // never shrinks
    std::vector<float> tab((size_t)len);
    const float cb = e->cfg.c_puct_base, ci = e->cfg.c_puct_init;
    for (long n = 0; n < len; ++n) tab[(size_t)n] = std::log(((float)(1 + (int)n) + cb) / cb) + ci;
    cudaFree(e->d_rate);
    OZ_TRY(dalloc(&e->d_rate, (size_t)len));
    OZ_CUDA(cudaMemcpy(e->d_rate, tab.data(), (size_t)len * sizeof(float), cudaMemcpyHostToDevice));
    e->t.rate = e->d_rate; e->t.rate_len = (int)len;
    return OZ_OK;
}

void sync_params(oz_engine *e) {
    Tree &t = e->t;
    t.eps = e->cfg.dirichlet_epsilon; t.alpha = e->cfg.dirichlet_alpha;
    t.seed_lo = (uint32_t)e->cfg.seed; t.seed_hi = (uint32_t)(e->cfg.seed >> 32);
    t.first_game_id = e->cfg.first_game_id; t.sym_mode = e->cfg.sym_mode; t.temperature_plies = e->cfg.temperature_plies;
    t.inv_temperature = 1.0f / e->cfg.self_play_temperature;
}

// The error bits travel to the pinned mirror with the stream; device_error_status() reads them after a synchronisation.
int fetch_device_errors(oz_engine *e) {
    OZ_CUDA(cudaMemcpyAsync(e->h_pin, e->t.error, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    return OZ_OK;
}
int device_error_status(oz_engine *e) {
    const int32_t err = e->h_pin[0];
    if (err == 0) return OZ_OK;
    cudaMemsetAsync(e->t.error, 0, sizeof(int32_t), e->stream);
    std::string msg = "device engine:";
    if (err & kErrArenaFull) msg += " node arena full (raise max_nodes_per_game);";
    if (err & kErrRateTable) msg += " visit count beyond the exploration-rate table;";
    if (err & kErrDepth) msg += " search path deeper than " + std::to_string(kMaxDepth) + ";";
    e->h_pin[0] = 0;
    return fail(OZ_ERR_CAPACITY, msg);
}
int check_device_errors(oz_engine *e) {
    OZ_TRY(fetch_device_errors(e));
    OZ_CUDA(cudaStreamSynchronize(e->stream));
    return device_error_status(e);
}

int set_device(const oz_engine *e) { OZ_CUDA(cudaSetDevice(e->cfg.device)); return OZ_OK; }

// Anything the captured round depends on (pointers, by-value parameters, evaluator) changed.
void drop_graph(oz_engine *e) {
    if (e->round_graph) { cudaGraphExecDestroy(e->round_graph); e->round_graph = nullptr; }
    e->warmed = false;
}

int grow_pools(oz_engine *e, int new_cap) {
    drop_graph(e);
    Tree nt = e->t;
    OZ_TRY(alloc_pools(nt, e->t.G, new_cap));
    launch_copy_arenas(e->t, nt, e->stream);
    OZ_CUDA(cudaStreamSynchronize(e->stream));
    free_pools(e->t);
    e->t = nt;
    return OZ_OK;
}

cudaEvent_t next_event(oz_engine *e, size_t i);
void read_profile(oz_engine *e);

// The reference's trees are heap-allocated and unbounded (search_node.h:29); ours live in fixed arenas, so before
// a search every game must have room for the worst case.  tree_bound is exact after a re-root (apply_kernel reports
// the largest tree) and grows by the worst case per search in between.
int ensure_capacity(oz_engine *e) {
    const long growth = search_growth(e->cfg);
    if (e->tree_bound + growth > e->t.cap) {           // tighten the bound before paying for a reallocation
        int32_t mx = 0;
        OZ_CUDA(cudaMemsetAsync(e->d_unfinished + 1, 0, sizeof(int32_t), e->stream));
        launch_max_count(e->t, e->d_unfinished + 1, e->stream);
        OZ_CUDA(cudaMemcpyAsync(&mx, e->d_unfinished + 1, sizeof(mx), cudaMemcpyDeviceToHost, e->stream));
        OZ_CUDA(cudaStreamSynchronize(e->stream));
        e->tree_bound = mx;
    }
    if (e->tree_bound + growth > e->t.cap) {
        const long want = std::max<long>(2L * e->t.cap, (e->tree_bound + growth) * 5 / 4);
        if (want > (1L << 24)) return fail(OZ_ERR_CAPACITY, "search tree beyond 16 M nodes per game");
        OZ_TRY(grow_pools(e, (int)want));
    }
    e->tree_bound += growth;
    return OZ_OK;
}

// The reference evaluates logf((1 + N + c_base) / c_base) for any N (search_thread.cpp:198-203) and allows any number
// of searches from the same root; the table indexed by N is therefore extended when the largest visit count a search
// can reach would leave it.  rootn_bound is a host-side bound (every search adds its simulations to the root; a child
// never has more visits than its parent); it is tightened with one read-back before paying for a rebuild.
int ensure_rate_table(oz_engine *e) {
    const long sims = (long)rounds_of(e->cfg) * e->cfg.num_threads * e->cfg.batch_size;
    if (e->rootn_bound + sims + 2 > e->t.rate_len) {
        std::vector<int32_t> rn((size_t)e->t.G);
        launch_root_stats(e->t, nullptr, nullptr, nullptr, e->d_stat_rootn, e->stream);
        OZ_CUDA(cudaMemcpyAsync(rn.data(), e->d_stat_rootn, rn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
        OZ_CUDA(cudaStreamSynchronize(e->stream));
        e->rootn_bound = *std::max_element(rn.begin(), rn.end());
    }
    if (e->rootn_bound + sims + 2 > e->t.rate_len) {
        if (e->rootn_bound + sims + 2 > (1L << 26)) return fail(OZ_ERR_CAPACITY, "visit count beyond 64 M: the exploration-rate table cannot grow further");
        drop_graph(e);
        OZ_CUDA(cudaStreamSynchronize(e->stream));
        OZ_TRY(build_rate_table(e, 2 * (e->rootn_bound + sims + 2)));
    }
    e->rootn_bound += sims;
    return OZ_OK;
}

// One evaluator call over `n` leaf slots starting at slot `first` (ws = network workspace = pipeline group).
// staged = the encode kernel already wrote the network's own input tensor (encode_target).
int eval_slots(oz_engine *e, size_t first, size_t n, cudaStream_t stream, int ws, cudaEvent_t conv_begin = nullptr, cudaEvent_t conv_end = nullptr,
               bool staged = false) {
    const int planes = 1 + 2 * e->t.H;
    const float *f = staged ? nullptr : e->d_features + first * planes * 64;
    float *p = e->d_policy + first * 65, *v = e->d_value + first;
    switch (e->cfg.evaluator) {
        case OZ_EVAL_SYNTH: launch_eval_synth(f, planes, n, p, v, stream); break;
        case OZ_EVAL_HASH: launch_eval_hash(f, planes, n, p, v, stream); break;
        case OZ_EVAL_RESNET:
            if (e->net.shape().in_planes != planes)
                return fail(OZ_ERR_STATE, "network in_planes " + std::to_string(e->net.shape().in_planes) + " != 1 + 2 * history_size");
            OZ_TRY(e->net.forward(f, n, p, v, stream, conv_begin, conv_end, ws));
            break;
        default: return fail(OZ_ERR_STATE, "oz_engine_search needs a built-in evaluator; use select/expand for OZ_EVAL_EXTERNAL");
    }
    e->ctr.evaluations += n;
    return OZ_OK;
}

// With the native network the encode kernel writes bf16 NHWC planes straight into the network's input tensor (the
// float32 planes of the reference's evaluator contract are only materialised for the other evaluators).
int encode_target(oz_engine *e, size_t n, int ws, void **x0) {
    *x0 = nullptr;
    if (e->cfg.evaluator != OZ_EVAL_RESNET) return OZ_OK;
    if (e->net.shape().in_planes != 1 + 2 * e->t.H)
        return fail(OZ_ERR_STATE, "network in_planes " + std::to_string(e->net.shape().in_planes) + " != 1 + 2 * history_size");
    return e->net.input_buffer(n, ws, x0);
}

cudaEvent_t next_event(oz_engine *e, size_t i) {
    while (e->ev.size() <= i) { cudaEvent_t x; cudaEventCreate(&x); e->ev.push_back(x); }
    return e->ev[i];
}

int begin_search(oz_engine *e) {
    if (e->in_search) return fail(OZ_ERR_STATE, "search already in progress");
    if (e->cfg.evaluator == OZ_EVAL_RESNET && !e->net.loaded()) return fail(OZ_ERR_STATE, "no network weights loaded");
    sync_params(e);
    OZ_TRY(ensure_capacity(e));
    OZ_TRY(ensure_rate_table(e));
    // New SearchThread per search: feature slots start from zeros (search_thread.cpp:37-39).  Only slots of terminal
    // leaves keep them, and their network outputs are never read, so the native network's own input tensor is not re-zeroed.
    if (e->cfg.evaluator != OZ_EVAL_RESNET)
        OZ_CUDA(cudaMemsetAsync(e->d_features, 0, (size_t)e->t.G * e->t.inflight * (1 + 2 * e->t.H) * 64 * sizeof(float), e->stream));
    OZ_CUDA(cudaMemsetAsync(e->t.round, 0, (1 + kMaxGroups) * sizeof(int32_t), e->stream));
    e->in_search = true; e->round = 0;
    return OZ_OK;
}

// `searched` = the games of this search (the whole tree or a range view), `active` = how many of them are unfinished.
int end_search(oz_engine *e, const Tree &searched, int active) {
    launch_end_search(searched, e->stream);
    e->in_search = false;
    e->ctr.searches += 1;
    e->ctr.simulations += (uint64_t)rounds_of(e->cfg) * e->t.inflight * (uint64_t)active;
    if (e->defer_error_check) return fetch_device_errors(e);     // play_ply: read after the ply's single synchronisation
    return check_device_errors(e);
}
int end_search(oz_engine *e) { return end_search(e, e->t, e->active); }

// ---- pipelined search ----
int pipeline_groups(const oz_engine *e) {
    if (e->cfg.evaluator == OZ_EVAL_EXTERNAL) return 1;
    int want = e->pipeline;
    // Automatic = off: measured on B200 (profiles/r01_pipeline_experiment.md) the convolutions lose more to the
    // co-running tree kernels (1 354 -> 1 160-1 210 TFLOP/s) than the hidden select / expand time gives back.
    if (want == 0) want = 1;
    return std::max(1, std::min({want, kMaxGroups, e->t.G}));
}

// The games [g0, g1) as a Tree of their own: every per-game, per-slot and per-node array is indexed by the local game.
Tree group_view(const oz_engine *e, int g0, int g1, int k) {
    Tree v = e->t;
    const size_t nb = (size_t)g0 * 2 * v.cap, sb = (size_t)g0 * v.inflight;
    v.n_p1 += nb; v.n_p2 += nb; v.n_meta += nb; v.n_parent += nb; v.n_child += nb; v.n_N += nb; v.n_W += nb; v.n_Q += nb; v.n_P += nb;
    v.g_root += g0; v.g_which += g0; v.g_count += g0; v.g_status += g0; v.g_ply += g0; v.g_search += g0; v.g_gameid += g0;
    v.h_p1 += (size_t)g0 * kMaxHist; v.h_p2 += (size_t)g0 * kMaxHist; v.h_player += (size_t)g0 * kMaxHist; v.h_n += g0;
    v.r_p1 += g0; v.r_p2 += g0; v.r_player += g0;
    v.s_leaf += sb; v.s_tf += sb; v.s_len += sb; v.s_path += sb * kMaxDepth;
    v.round = e->t.round + 1 + k;
    v.G = g1 - g0;
    return v;
}

// Length of the union of [a, b) intervals.
double union_ms(std::vector<std::pair<float, float>> iv) {
    std::sort(iv.begin(), iv.end());
    double total = 0.0; float hi = -1e30f;
    for (const auto &x : iv) {
        if (x.second <= hi) continue;
        total += (double)x.second - (double)std::max(x.first, hi);
        hi = x.second;
    }
    return total;
}

// Phase times of the last search from its CUDA events (which must have completed: called after a synchronisation).
void read_profile(oz_engine *e) {
    const size_t n_ev = e->prof_events;
    e->prof_events = 0;
    if (n_ev == 0) return;
    const bool net_prof = e->prof_net;
    if (e->prof_pipelined) {
        // Phases of different groups overlap: the network and the convolutions are charged the time at least one group
        // spent in them (union of intervals on the fork event's clock); select / expand are the per-group sums.
        std::vector<std::pair<float, float>> net_iv, conv_iv;
        auto at = [&](size_t i) { float ms = 0.f; cudaEventElapsedTime(&ms, e->ev_fork, e->ev[i]); return ms; };
        for (size_t i = 0; i + 5 < n_ev; i += 6) {
            const float t0 = at(i), t1 = at(i + 1), t4 = at(i + 4), t5 = at(i + 5);
            e->ctr.select_ms += std::max(0.f, t1 - t0); e->ctr.expand_ms += std::max(0.f, t5 - t4);
            net_iv.emplace_back(t1, t4);
            if (net_prof) { conv_iv.emplace_back(at(i + 2), at(i + 3)); e->ctr.conv_launches += (uint64_t)(1 + 2 * e->net.shape().blocks); }
        }
        e->ctr.eval_ms += union_ms(net_iv);
        e->ctr.conv_ms += union_ms(conv_iv);
        return;
    }
    for (size_t i = 0; i + 5 < n_ev; i += 6) {
        float a = 0, b = 0, c = 0, d = 0;
        cudaEventElapsedTime(&a, e->ev[i], e->ev[i + 1]); cudaEventElapsedTime(&b, e->ev[i + 1], e->ev[i + 4]); cudaEventElapsedTime(&c, e->ev[i + 4], e->ev[i + 5]);
        e->ctr.select_ms += a; e->ctr.eval_ms += b; e->ctr.expand_ms += c;
        if (net_prof) { cudaEventElapsedTime(&d, e->ev[i + 2], e->ev[i + 3]); e->ctr.conv_ms += d; e->ctr.conv_launches += (uint64_t)(1 + 2 * e->net.shape().blocks); }
    }
}

int search_pipelined(oz_engine *e, int groups) {
    const int rounds = rounds_of(e->cfg), G = e->t.G, B = e->t.inflight;
    // Tree streams and the network stream get the SAME priority.  Round 1 put the network on a high-priority stream; with
    // programmatic dependent launch there is then always a pending, not yet placeable high-priority convolution grid, and the
    // block scheduler does not place lower-priority blocks while it waits: nothing overlapped (profiles/r02_dynamic_scheduler.md).
    // OZ_PIPE_PRIO=0 restores that setting, =2 inverts it (experiment switches).
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    {
        const char *pe = getenv("OZ_PIPE_PRIO");
        const int mode = pe ? atoi(pe) : 1;
        if (mode == 1) prio_hi = prio_lo;
        else if (mode == 2) std::swap(prio_lo, prio_hi);
    }
    for (int k = 0; k < groups; ++k) {
        if (!e->gstream[k] && cudaStreamCreateWithPriority(&e->gstream[k], cudaStreamNonBlocking, prio_lo) != cudaSuccess) { e->in_search = false; return fail(OZ_ERR_CUDA, "stream creation failed"); }
        if (!e->ev_join[k]) { cudaEventCreateWithFlags(&e->ev_join[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&e->ev_sel[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&e->ev_net[k], cudaEventDisableTiming); }
    }
    if (!e->nstream && cudaStreamCreateWithPriority(&e->nstream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { e->in_search = false; return fail(OZ_ERR_CUDA, "stream creation failed"); }
    if (!e->ev_fork) cudaEventCreate(&e->ev_fork);
    const bool net_prof = e->profiling && e->cfg.evaluator == OZ_EVAL_RESNET;
    cudaEventRecord(e->ev_fork, e->stream);
    for (int k = 0; k < groups; ++k) cudaStreamWaitEvent(e->gstream[k], e->ev_fork, 0);
    cudaStreamWaitEvent(e->nstream, e->ev_fork, 0);
    e->net.set_pdl(e->pipeline_pdl);
    size_t n_ev = 0;
    int st = OZ_OK;
    // OZ_PIPE_WIDE=1 (experiment switch): the wide tree kernels, which do not fit beside a convolution CTA and therefore run
    // only on SMs the convolution leaves free (OZ_CONV_CLUSTERS), instead of the narrow co-resident ones.
    static int wide = -1;
    if (wide < 0) { const char *we = getenv("OZ_PIPE_WIDE"); wide = (we && we[0] == '1') ? 1 : 0; }
    const bool narrow = wide == 0;
    NvtxRange nvtx_pipe("oz: pipelined rounds");
    for (int r = 0; r < rounds && !st; ++r) {
        e->round = r;
        for (int k = 0; k < groups && !st; ++k) {
            const int g0 = (int)((long)G * k / groups), g1 = (int)((long)G * (k + 1) / groups);
            const Tree v = group_view(e, g0, g1, k);
            cudaStream_t ts = e->gstream[k], ns = e->nstream;
            const size_t first = (size_t)g0 * B, n = (size_t)(g1 - g0) * B;
            cudaEvent_t cb = nullptr, ce = nullptr;
            if (e->profiling) { cudaEventRecord(next_event(e, n_ev), ts); cb = next_event(e, n_ev + 2); ce = next_event(e, n_ev + 3); }
            void *x0 = nullptr;
            if ((st = encode_target(e, n, k, &x0))) break;
            launch_select(v, ts, narrow);
            launch_encode(v, e->d_features + first * (1 + 2 * e->t.H) * 64, ts, narrow, x0);
            cudaEventRecord(e->ev_sel[k], ts);
            cudaStreamWaitEvent(ns, e->ev_sel[k], 0);
            if (e->profiling) cudaEventRecord(next_event(e, n_ev + 1), ns);
            st = eval_slots(e, first, n, ns, k, net_prof ? cb : nullptr, net_prof ? ce : nullptr, x0 != nullptr);
            if (e->profiling) cudaEventRecord(next_event(e, n_ev + 4), ns);
            cudaEventRecord(e->ev_net[k], ns);
            cudaStreamWaitEvent(ts, e->ev_net[k], 0);
            if (!st) launch_expand(v, e->d_policy + first * 65, e->d_value + first, ts, narrow);
            if (e->profiling) { cudaEventRecord(next_event(e, n_ev + 5), ts); n_ev += 6; }
        }
    }
    e->net.set_pdl(true);
    for (int k = 0; k < groups; ++k) { cudaEventRecord(e->ev_join[k], e->gstream[k]); cudaStreamWaitEvent(e->stream, e->ev_join[k], 0); }
    if (st) { cudaStreamSynchronize(e->stream); e->in_search = false; return st; }
    e->round = rounds;
    e->warmed = true;
    st = end_search(e);
    if (!st && e->profiling) { e->prof_events = n_ev; e->prof_pipelined = true; e->prof_net = net_prof; if (!e->defer_error_check) read_profile(e); }
    return st;
}

}  // namespace

extern "C" {

void oz_config_default(oz_config *c) {
    std::memset(c, 0, sizeof(*c));
    c->num_games = 1; c->history_size = 4; c->num_simulations = 800; c->num_threads = 2; c->batch_size = 16;
    c->c_puct_base = 20000.0f; c->c_puct_init = 2.5f; c->dirichlet_epsilon = 0.25f; c->dirichlet_alpha = 0.5f;
    c->seed = 0; c->first_game_id = 0; c->device = 0; c->evaluator = OZ_EVAL_EXTERNAL; c->sym_mode = OZ_SYM_RANDOM;
    c->temperature_plies = 12; c->max_nodes_per_game = 0; c->self_play_temperature = 1.0f; c->reserved = 0;
}

int oz_engine_create(const oz_config *cfg, oz_engine **out) {
    *out = nullptr;
    OZ_TRY(check_config(*cfg));
    OZ_TRY(check_engine_limits(*cfg));
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(OZ_ERR_CUDA, "no CUDA device: the engine has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(OZ_ERR_CUDA, "invalid CUDA device " + std::to_string(cfg->device));
    oz_engine *e = new oz_engine();
    e->cfg = *cfg;
    int st = OZ_OK;
    do {
        if ((st = set_device(e))) break;
        if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { st = fail(OZ_ERR_CUDA, "stream creation failed"); break; }
        Tree &t = e->t;
        const int G = cfg->num_games;
        t.G = G;
        if ((st = alloc_pools(t, G, default_cap(*cfg)))) break;
        if ((st = dalloc(&t.g_root, G)) || (st = dalloc(&t.g_which, G)) || (st = dalloc(&t.g_count, G)) || (st = dalloc(&t.g_status, G)) ||
            (st = dalloc(&t.g_ply, G)) || (st = dalloc(&t.g_search, G)) || (st = dalloc(&t.g_gameid, G)) || (st = dalloc(&t.h_p1, (size_t)G * kMaxHist)) ||
            (st = dalloc(&t.h_p2, (size_t)G * kMaxHist)) || (st = dalloc(&t.h_player, (size_t)G * kMaxHist)) || (st = dalloc(&t.h_n, G)) ||
            (st = dalloc(&t.r_p1, G)) || (st = dalloc(&t.r_p2, G)) || (st = dalloc(&t.r_player, G)) || (st = dalloc(&t.error, 1)) ||
            (st = dalloc(&t.stats, 4 * kStatCopies)) || (st = dalloc(&t.round, 1 + kMaxGroups)) || (st = dalloc(&e->d_actions, G)) || (st = dalloc(&e->d_unfinished, 2)) ||
            (st = dalloc(&e->d_stat_n, G)) || (st = dalloc(&e->d_stat_counts, (size_t)G * kStride)) || (st = dalloc(&e->d_stat_q, (size_t)G * kStride)) ||
            (st = dalloc(&e->d_stat_rootn, G)) || (st = dalloc(&e->d_spd_p, 8 * 65)))
            break;
        cudaMemset(t.error, 0, sizeof(int32_t)); cudaMemset(t.stats, 0, 4 * kStatCopies * sizeof(unsigned long long));
        e->max_plies = 128;
        if ((st = dalloc(&e->d_records, (size_t)G * e->max_plies))) break;
        if (cudaHostAlloc((void **)&e->h_pin, 8 * sizeof(int32_t), cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc((void **)&e->h_actions, (size_t)G * sizeof(int32_t), cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc((void **)&e->h_counts, (size_t)G * kStride * sizeof(int32_t), cudaHostAllocDefault) != cudaSuccess) {
            st = fail(OZ_ERR_CUDA, "pinned host allocation failed"); break;
        }
        std::memset(e->h_pin, 0, 8 * sizeof(int32_t));
        if ((st = alloc_slots(e))) break;
        if ((st = build_rate_table(e))) break;
        sync_params(e);
        launch_reset(t, e->stream);
        e->active = G;
        if (cudaStreamSynchronize(e->stream) != cudaSuccess) { st = fail(OZ_ERR_CUDA, std::string("engine init: ") + cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    if (st) { oz_engine_destroy(e); return st; }
    *out = e;
    return OZ_OK;
}

void oz_engine_destroy(oz_engine *e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    Tree &t = e->t;
    free_pools(t);
    drop_graph(e);
    void *ptrs[] = {t.round, t.g_gameid, e->done.records, e->done.plies, e->done.gameid, e->done.p1, e->done.p2, e->done.counters, e->done.slot_done,
                    e->done.slot_new, t.g_root, t.g_which, t.g_count, t.g_status, t.g_ply, t.g_search, t.h_p1, t.h_p2, t.h_player, t.h_n, t.r_p1, t.r_p2,
                    t.r_player, t.s_leaf, t.s_tf, t.s_len, t.s_path, t.error, t.stats, e->d_features, e->d_policy, e->d_value, e->d_rate,
                    e->d_gamma, e->d_actions, e->d_unfinished, e->d_stat_n, e->d_stat_counts, e->d_stat_q, e->d_stat_rootn, e->d_records,
                    e->d_spd_f, e->d_spd_p};
    for (void *p : ptrs) cudaFree(p);
    cudaFreeHost(e->h_pin); cudaFreeHost(e->h_actions); cudaFreeHost(e->h_counts);
    for (cudaEvent_t x : e->ev) cudaEventDestroy(x);
    for (int k = 0; k < kMaxGroups; ++k) {
        if (e->gstream[k]) cudaStreamDestroy(e->gstream[k]);
        if (e->ev_join[k]) { cudaEventDestroy(e->ev_join[k]); cudaEventDestroy(e->ev_sel[k]); cudaEventDestroy(e->ev_net[k]); }
    }
    if (e->nstream) cudaStreamDestroy(e->nstream);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

#define OZ_SETTER(NAME, FIELD, TYPE, AFTER)                      \
    int oz_engine_set_##NAME(oz_engine *e, TYPE v) {             \
        oz_config c = e->cfg;                                    \
        c.FIELD = v;                                             \
        OZ_TRY(check_config(c));                                 \
        OZ_TRY(check_engine_limits(c));                          \
        if (e->in_search) return fail(OZ_ERR_STATE, "cannot change parameters during a search"); \
        OZ_TRY(set_device(e));                                   \
        drop_graph(e);                                           \
        e->cfg = c;                                              \
        AFTER;                                                   \
        return OZ_OK;                                            \
    }
OZ_SETTER(history_size, history_size, int, OZ_TRY(alloc_slots(e)))
OZ_SETTER(num_simulations, num_simulations, int, { OZ_TRY(build_rate_table(e)); if (default_cap(e->cfg) > e->t.cap) OZ_TRY(grow_pools(e, default_cap(e->cfg))); })
OZ_SETTER(num_threads, num_threads, int, { OZ_TRY(alloc_slots(e)); OZ_TRY(build_rate_table(e)); })
OZ_SETTER(batch_size, batch_size, int, { OZ_TRY(alloc_slots(e)); OZ_TRY(build_rate_table(e)); })
OZ_SETTER(c_puct_base, c_puct_base, float, OZ_TRY(build_rate_table(e)))
OZ_SETTER(c_puct_init, c_puct_init, float, OZ_TRY(build_rate_table(e)))
OZ_SETTER(dirichlet_epsilon, dirichlet_epsilon, float, sync_params(e))
OZ_SETTER(dirichlet_alpha, dirichlet_alpha, float, sync_params(e))

int oz_engine_get_config(const oz_engine *e, oz_config *out) { *out = e->cfg; return OZ_OK; }

int oz_engine_set_evaluator(oz_engine *e, int evaluator) {
    if (evaluator < OZ_EVAL_EXTERNAL || evaluator > OZ_EVAL_RESNET) return fail(OZ_ERR_INVALID_ARGUMENT, "unknown evaluator");
    if (e->in_search) return fail(OZ_ERR_STATE, "cannot change the evaluator during a search");
    if (e->cfg.evaluator != evaluator) drop_graph(e);
    e->cfg.evaluator = evaluator;
    return OZ_OK;
}

int oz_engine_set_gamma_table(oz_engine *e, const float *table, int len) {
    OZ_TRY(set_device(e));
    drop_graph(e);
    cudaFree(e->d_gamma); e->d_gamma = nullptr;
    e->t.gamma_table = nullptr; e->t.gamma_len = 0;
    if (len > 0) {
        OZ_TRY(dalloc(&e->d_gamma, (size_t)len));
        OZ_CUDA(cudaMemcpy(e->d_gamma, table, (size_t)len * sizeof(float), cudaMemcpyHostToDevice));
        e->t.gamma_table = e->d_gamma; e->t.gamma_len = len;
    }
    return OZ_OK;
}

int oz_engine_reset(oz_engine *e) {
    OZ_TRY(set_device(e));
    if (e->in_search) return fail(OZ_ERR_STATE, "cannot reset during a search");
    sync_params(e);
    launch_reset(e->t, e->stream);
    e->active = e->t.G;
    e->tree_bound = 1;
    e->rootn_bound = 0;
    if (e->done.capacity > 0) {
        const int32_t init[3] = {0, e->t.G, 0};
        OZ_CUDA(cudaMemcpyAsync(e->done.counters, init, sizeof(init), cudaMemcpyHostToDevice, e->stream));
    }
    OZ_CUDA(cudaStreamSynchronize(e->stream));
    return OZ_OK;
}

int oz_engine_set_game_pool(oz_engine *e, int total_games) {
    OZ_TRY(set_device(e));
    if (e->in_search) return fail(OZ_ERR_STATE, "cannot change the game pool during a search");
    DoneStore &d = e->done;
    cudaFree(d.records); cudaFree(d.plies); cudaFree(d.gameid); cudaFree(d.p1); cudaFree(d.p2); cudaFree(d.counters);
    cudaFree(d.slot_done); cudaFree(d.slot_new);
    d = DoneStore{};
    e->pool_total = 0;
    if (total_games == 0) return oz_engine_reset(e);
    if (total_games < e->t.G)
        return fail(OZ_ERR_INVALID_ARGUMENT, "Expected total_games >= num_games, but got " + std::to_string(total_games) + ".");
    OZ_TRY(dalloc(&d.records, (size_t)total_games * e->max_plies)); OZ_TRY(dalloc(&d.plies, total_games));
    OZ_TRY(dalloc(&d.gameid, total_games)); OZ_TRY(dalloc(&d.p1, total_games)); OZ_TRY(dalloc(&d.p2, total_games));
    OZ_TRY(dalloc(&d.counters, 4)); OZ_TRY(dalloc(&d.slot_done, e->t.G)); OZ_TRY(dalloc(&d.slot_new, e->t.G));
    d.capacity = total_games;
    e->pool_total = total_games;
    return oz_engine_reset(e);
}

int oz_engine_finished_games(oz_engine *e, int32_t *count) {
    OZ_TRY(set_device(e));
    if (e->done.capacity == 0) return fail(OZ_ERR_STATE, "no game pool: call oz_engine_set_game_pool first");
    OZ_CUDA(cudaMemcpy(count, e->done.counters, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return OZ_OK;
}

int oz_engine_finished_records(oz_engine *e, uint32_t *game_ids, int32_t *plies, oz_ply_record *records, int max_plies, int capacity,
                               int32_t *count) {
    OZ_TRY(oz_engine_finished_games(e, count));
    const int n = std::min<int>(*count, capacity);
    if (n <= 0) return OZ_OK;
    if (game_ids) OZ_CUDA(cudaMemcpy(game_ids, e->done.gameid, (size_t)n * 4, cudaMemcpyDeviceToHost));
    std::vector<int32_t> pl(n);
    OZ_CUDA(cudaMemcpy(pl.data(), e->done.plies, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (plies) std::copy(pl.begin(), pl.end(), plies);
    if (records) {
        if (max_plies == e->max_plies) {
            OZ_CUDA(cudaMemcpy(records, e->done.records, (size_t)n * max_plies * sizeof(oz_ply_record), cudaMemcpyDeviceToHost));
        } else {
            for (int i = 0; i < n; ++i) {
                const int k = std::min(pl[i], max_plies);
                if (k > 0) OZ_CUDA(cudaMemcpy(records + (size_t)i * max_plies, e->done.records + (size_t)i * e->max_plies, (size_t)k * sizeof(oz_ply_record), cudaMemcpyDeviceToHost));
            }
        }
    }
    return OZ_OK;
}

int oz_engine_positions(oz_engine *e, oz_position *out) {
    OZ_TRY(set_device(e));
    const int G = e->t.G;
    std::vector<uint64_t> p1(G), p2(G);
    std::vector<int32_t> pl(G);
    OZ_CUDA(cudaMemcpyAsync(p1.data(), e->t.r_p1, G * 8, cudaMemcpyDeviceToHost, e->stream));
    OZ_CUDA(cudaMemcpyAsync(p2.data(), e->t.r_p2, G * 8, cudaMemcpyDeviceToHost, e->stream));
    OZ_CUDA(cudaMemcpyAsync(pl.data(), e->t.r_player, G * 4, cudaMemcpyDeviceToHost, e->stream));
    OZ_CUDA(cudaStreamSynchronize(e->stream));
    for (int g = 0; g < G; ++g) {
        // (player, p1, p2) determine the move masks (position.h:349-358).
        oz_position &o = out[g];
        o.player = pl[g]; o.p1 = p1[g]; o.p2 = p2[g]; o.legal = 0; o.next_legal = 0;
        if (pl[g] != 0) {
            const uint64_t me = pl[g] == 1 ? p1[g] : p2[g], opp = pl[g] == 1 ? p2[g] : p1[g];
            o.legal = legal_moves(me, opp);
            if (o.legal == 0) o.next_legal = legal_moves(opp, me);
        }
    }
    return OZ_OK;
}

size_t oz_resnet_blob_bytes(const oz_resnet_shape *s) { return resnet_layout(*s).head_w; }

int oz_resnet_blob_offsets(const oz_resnet_shape *s, int64_t *offsets, int capacity, int *count) {
    ResNetLayout L = resnet_layout(*s);
    std::vector<int64_t> v;
    for (size_t i = 0; i < L.conv_w.size(); ++i) { v.push_back((int64_t)L.conv_w[i]); v.push_back((int64_t)L.conv_b[i]); }
    for (size_t o : {L.pconv_w, L.pconv_b, L.pfc_w, L.pfc_b, L.vconv_w, L.vconv_b, L.vfc1_w, L.vfc1_b, L.vfc2_w, L.vfc2_b}) v.push_back((int64_t)o);
    *count = (int)v.size();
    if (capacity < (int)v.size()) return fail(OZ_ERR_INVALID_ARGUMENT, "offset buffer too small");
    std::copy(v.begin(), v.end(), offsets);
    return OZ_OK;
}

int oz_engine_load_resnet(oz_engine *e, const oz_resnet_shape *shape, const void *blob, size_t bytes) {
    OZ_TRY(set_device(e));
    if (e->in_search) return fail(OZ_ERR_STATE, "cannot load weights during a search");
    OZ_CUDA(cudaStreamSynchronize(e->stream));
    drop_graph(e);
    OZ_TRY(e->net.load(*shape, blob, bytes));
    OZ_TRY(e->net.reserve((size_t)e->t.G * e->t.inflight));
    return OZ_OK;
}

// ---- the network alone ----
int oz_resnet_create(int device, oz_resnet **out) {
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(OZ_ERR_CUDA, "no such CUDA device");
    oz_resnet *n = new oz_resnet();
    n->device = device;
    *out = n;
    return OZ_OK;
}
void oz_resnet_destroy(oz_resnet *n) {
    if (!n) return;
    cudaSetDevice(n->device);
    cudaFree(n->d_f); cudaFree(n->d_p); cudaFree(n->d_v);
    delete n;
}
int oz_resnet_load(oz_resnet *n, const oz_resnet_shape *shape, const void *blob, size_t bytes) {
    OZ_CUDA(cudaSetDevice(n->device));
    return n->net.load(*shape, blob, bytes);
}
int oz_resnet_forward_device(oz_resnet *n, const float *f, size_t cnt, float *p, float *v, void *stream) {
    OZ_CUDA(cudaSetDevice(n->device));
    return n->net.forward(f, cnt, p, v, (cudaStream_t)stream);
}
int oz_resnet_forward(oz_resnet *n, const float *f, size_t cnt, float *p, float *v) {
    OZ_CUDA(cudaSetDevice(n->device));
    if (!n->net.loaded()) return fail(OZ_ERR_STATE, "no network weights loaded");
    if (cnt == 0) return OZ_OK;
    const size_t planes = n->net.shape().in_planes;
    if (cnt > n->cap) {
        cudaFree(n->d_f); cudaFree(n->d_p); cudaFree(n->d_v);
        OZ_TRY(dalloc(&n->d_f, cnt * planes * 64)); OZ_TRY(dalloc(&n->d_p, cnt * 65)); OZ_TRY(dalloc(&n->d_v, cnt));
        n->cap = cnt;
    }
    OZ_CUDA(cudaMemcpy(n->d_f, f, cnt * planes * 64 * sizeof(float), cudaMemcpyHostToDevice));
    OZ_TRY(n->net.forward(n->d_f, cnt, n->d_p, n->d_v, nullptr));
    OZ_CUDA(cudaMemcpy(p, n->d_p, cnt * 65 * sizeof(float), cudaMemcpyDeviceToHost));
    OZ_CUDA(cudaMemcpy(v, n->d_v, cnt * sizeof(float), cudaMemcpyDeviceToHost));
    return OZ_OK;
}

int oz_debug_set_tower(int mode) {
    oz::set_tower_mode(mode);
    return OZ_OK;
}
int oz_debug_set_heads_boards(int boards) {
    oz::set_heads_boards(boards);
    return OZ_OK;
}
int oz_debug_tower_timeline(long long *out, int *max_clusters) {
    if (max_clusters) *max_clusters = oz::tower_max_clusters();
    if (!out) return OZ_OK;
    if (!oz::tower_dbg_ptr()) return fail(OZ_ERR_STATE, "no tower timeline: set OZ_TOWER_DBG=0 in the environment before the first forward");
    OZ_CUDA(cudaDeviceSynchronize());
    OZ_CUDA(cudaMemcpy(out, oz::tower_dbg_ptr(), 64 * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    return OZ_OK;
}

int oz_debug_conv3x3(const uint16_t *in, int cin, const uint16_t *w, const float *bias, const uint16_t *skip, int relu,
                     uint16_t *out, int cout, size_t boards, double *device_ms) {
    const size_t pad = (boards + kBoardPad - 1) / kBoardPad * kBoardPad;
    uint16_t *d_in = nullptr, *d_w = nullptr, *d_skip = nullptr, *d_out = nullptr;
    float *d_b = nullptr;
    int *d_work = nullptr;
    int st = OZ_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        if ((st = dalloc(&d_in, pad * 64 * cin)) || (st = dalloc(&d_w, (size_t)9 * cout * cin)) || (st = dalloc(&d_b, (size_t)cout)) ||
            (st = dalloc(&d_out, pad * 64 * cout)) || (skip && (st = dalloc(&d_skip, pad * 64 * cout))) || (st = dalloc(&d_work, 1))) break;
        cudaMemset(d_in, 0, pad * 64 * cin * 2); cudaMemset(d_out, 0, pad * 64 * cout * 2);
        if (skip) { cudaMemset(d_skip, 0, pad * 64 * cout * 2); cudaMemcpy(d_skip, skip, boards * 64 * cout * 2, cudaMemcpyHostToDevice); }
        cudaMemcpy(d_in, in, boards * 64 * cin * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(d_w, w, (size_t)9 * cout * cin * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(d_b, bias, cout * sizeof(float), cudaMemcpyHostToDevice);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if ((st = ResNet::conv3x3(d_in, cin, d_w, d_b, d_skip, relu, d_out, cout, pad, nullptr, d_work))) break;
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { st = fail(OZ_ERR_CUDA, std::string("conv3x3: ") + cudaGetErrorString(err)); break; }
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (device_ms) *device_ms = ms;
        cudaMemcpy(out, d_out, boards * 64 * cout * 2, cudaMemcpyDeviceToHost);
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_in); cudaFree(d_w); cudaFree(d_skip); cudaFree(d_out); cudaFree(d_b); cudaFree(d_work);
    return st;
}

int oz_debug_wgrad3x3(const uint16_t *x, int cinp, const uint16_t *dz, int cout, size_t boards, float *G, double *device_ms) {
    uint16_t *d_x = nullptr, *d_dz = nullptr;
    float *d_G = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int st = OZ_OK;
    do {
        if ((st = dalloc(&d_x, boards * 64 * cinp)) || (st = dalloc(&d_dz, boards * 64 * cout)) || (st = dalloc(&d_G, (size_t)cout * 9 * cinp))) break;
        cudaMemcpy(d_x, x, boards * 64 * cinp * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(d_dz, dz, boards * 64 * cout * 2, cudaMemcpyHostToDevice);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if ((st = launch_wgrad3x3(d_x, cinp, d_dz, cout, boards, d_G, nullptr))) break;
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { st = fail(OZ_ERR_CUDA, std::string("wgrad3x3: ") + cudaGetErrorString(err)); break; }
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (device_ms) *device_ms = ms;
        cudaMemcpy(G, d_G, (size_t)cout * 9 * cinp * sizeof(float), cudaMemcpyDeviceToHost);
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_x); cudaFree(d_dz); cudaFree(d_G);
    return st;
}

// ---- search ----
int oz_engine_rounds(const oz_engine *e) { return rounds_of(e->cfg); }

int oz_engine_search_begin(oz_engine *e) {
    OZ_TRY(set_device(e));
    OZ_TRY(begin_search(e));
    OZ_CUDA(cudaStreamSynchronize(e->stream));   // the caller's stream takes over from here
    return OZ_OK;
}

int oz_engine_select_device(oz_engine *e, float *features, void *stream) {
    if (!e->in_search) return fail(OZ_ERR_STATE, "search_begin not called");
    OZ_TRY(set_device(e));
    cudaStream_t s = (cudaStream_t)stream;    // NULL is the legacy default stream, as in the CUDA runtime
    launch_select(e->t, s);
    launch_encode(e->t, features ? features : e->d_features, s);
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

int oz_engine_expand_device(oz_engine *e, const float *policy, const float *value, void *stream) {
    if (!e->in_search) return fail(OZ_ERR_STATE, "search_begin not called");
    OZ_TRY(set_device(e));
    cudaStream_t s = (cudaStream_t)stream;
    launch_expand(e->t, policy, value, s);
    e->round += 1;
    e->ctr.evaluations += (uint64_t)e->t.G * e->t.inflight;
    OZ_CUDA(cudaGetLastError());
    return OZ_OK;
}

int oz_engine_search_end(oz_engine *e) {
    if (!e->in_search) return fail(OZ_ERR_STATE, "search_begin not called");
    OZ_TRY(set_device(e));
    OZ_CUDA(cudaDeviceSynchronize());   // external evaluators may have used their own streams
    return end_search(e);
}

// Unfinished games among [g0, g1) (one small read-back; only range searches need it).
static int count_unfinished(oz_engine *e, int g0, int g1, int *out) {
    std::vector<int32_t> st((size_t)(g1 - g0));
    OZ_CUDA(cudaMemcpyAsync(st.data(), e->t.g_status + g0, st.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    OZ_CUDA(cudaStreamSynchronize(e->stream));
    int n = 0;
    for (int32_t x : st) n += (x & 1) ? 0 : 1;
    *out = n;
    return OZ_OK;
}

static int search_games(oz_engine *e, int g0, int g1) {
    NvtxRange nvtx_search("oz: search");
    OZ_TRY(set_device(e));
    const bool whole = g0 == 0 && g1 == e->t.G;
    int active = e->active;
    if (!whole) OZ_TRY(count_unfinished(e, g0, g1, &active));
    OZ_TRY(begin_search(e));
    const int rounds = rounds_of(e->cfg);
    size_t k = 0;
    int st = OZ_OK;
    const bool net_prof = e->profiling && e->cfg.evaluator == OZ_EVAL_RESNET;
    // Launch-bound batches (a single game is 26 launches of a few microseconds per round) replay one captured round;
    // the round index lives on the device, so the same graph serves every round of every search.
    const size_t boards = (size_t)e->t.G * e->t.inflight;
    const bool want_graph = whole && !e->profiling && e->warmed && (e->graph_mode == 2 || (e->graph_mode == 1 && boards <= 8192));
    const int groups = whole ? pipeline_groups(e) : 1;
    if (groups > 1 && !want_graph) return search_pipelined(e, groups);
    // Graph-replayed searches split the games into BRANCHES instead: game groups are independent for a whole search, so each
    // group's rounds are a chain of their own (own stream while capturing, own network workspace, own round counter) and the
    // branches run side by side -- one group's tree kernels and kernel boundaries under another group's MMAs.
    // Automatic (pipeline = 0): on where it was measured to pay -- a network the cluster tower serves, more 32-board clusters
    // than the device runs at once (7) and at most 48: one tower wave per branch instead of waves of one launch, or instead of
    // the per-layer path (128x10b, 8 / 16 / 24 / 32 games: 8.6 -> 6.7, 10.4 -> 7.2, 12.1 -> 8.3, 12.9 -> 10.3 ms per ply;
    // no change below 8 games, slower from 64 games; 256-channel networks: slower at every size).
    int want_branches = e->pipeline > 1 ? e->pipeline : 1;
    if (e->pipeline == 0 && e->cfg.evaluator == OZ_EVAL_RESNET && e->net.tower_capable()) {
        const size_t clusters = (boards + 31) / 32;
        if (clusters > 7 && clusters <= 48) want_branches = kMaxGroups;
    }
    const int branches = want_graph ? std::max(1, std::min({want_branches, kMaxGroups, e->t.G})) : 1;
    if (want_graph && e->round_graph && (e->graph_rounds != rounds || e->graph_branches != branches)) drop_graph(e);   // num_simulations / batch / pipeline changed
    if (want_graph && !e->round_graph) {
        cudaGraph_t g = nullptr;
        const unsigned long long before = g_kernel_launches;
        const uint64_t evals_before = e->ctr.evaluations;
        void *gx0[kMaxGroups] = {};
        int bg0[kMaxGroups + 1] = {};
        for (int k = 0; k <= branches; ++k) bg0[k] = (int)((long)e->t.G * k / branches);
        for (int k = 0; k < branches && !st; ++k) {      // everything that allocates happens before the capture
            st = encode_target(e, (size_t)(bg0[k + 1] - bg0[k]) * e->t.inflight, k, &gx0[k]);
            if (!st && branches > 1) {
                if (!e->gstream[k] && cudaStreamCreateWithFlags(&e->gstream[k], cudaStreamNonBlocking) != cudaSuccess) st = fail(OZ_ERR_CUDA, "stream creation failed");
                if (!e->ev_join[k]) { cudaEventCreateWithFlags(&e->ev_join[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&e->ev_sel[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&e->ev_net[k], cudaEventDisableTiming); }
            }
        }
        if (!st && branches > 1 && !e->ev_fork) cudaEventCreate(&e->ev_fork);
        if (st) { e->in_search = false; return st; }
        if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { e->in_search = false; return fail(OZ_ERR_CUDA, "cudaStreamBeginCapture failed"); }
        // Every round of the search in ONE graph, the kernels chained by programmatic dependent launch: each becomes resident
        // while its predecessor runs and waits (griddepcontrol.wait) before it touches the predecessor's results.
        if (branches > 1) {
            cudaEventRecord(e->ev_fork, e->stream);
            for (int k = 0; k < branches; ++k) cudaStreamWaitEvent(e->gstream[k], e->ev_fork, 0);
        }
        for (int r = 0; r < rounds && !st; ++r)
            for (int k = 0; k < branches && !st; ++k) {
                const Tree v = branches > 1 ? group_view(e, bg0[k], bg0[k + 1], k) : e->t;
                cudaStream_t bs = branches > 1 ? e->gstream[k] : e->stream;
                const size_t first = (size_t)bg0[k] * e->t.inflight, n = (size_t)(bg0[k + 1] - bg0[k]) * e->t.inflight;
                launch_select(v, bs, false, true);
                launch_encode(v, e->d_features + first * (1 + 2 * e->t.H) * 64, bs, false, gx0[k], true);
                st = eval_slots(e, first, n, bs, k, nullptr, nullptr, gx0[k] != nullptr);
                if (!st) launch_expand(v, e->d_policy + first * 65, e->d_value + first, bs, false, true);
            }
        if (branches > 1)
            for (int k = 0; k < branches; ++k) { cudaEventRecord(e->ev_join[k], e->gstream[k]); cudaStreamWaitEvent(e->stream, e->ev_join[k], 0); }
        e->graph_rounds = rounds; e->graph_branches = branches;
        cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        e->graph_launches = g_kernel_launches - before;
        g_kernel_launches = before;                    // capturing launched nothing
        e->ctr.evaluations = evals_before;
        if (st || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            e->in_search = false;
            return st ? st : fail(OZ_ERR_CUDA, std::string("graph capture of a search round failed: ") + cudaGetErrorString(ce));
        }
        ce = cudaGraphInstantiate(&e->round_graph, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) { e->round_graph = nullptr; e->in_search = false; return fail(OZ_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
    }
    if (want_graph) {
        const cudaError_t ge = cudaGraphLaunch(e->round_graph, e->stream);
        if (ge != cudaSuccess) { e->in_search = false; return fail(OZ_ERR_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(ge)); }
        count_launch((int)e->graph_launches);
        e->ctr.evaluations += boards * (uint64_t)rounds;
        e->round = rounds;
        return end_search(e);
    }
    // A range search walks a view of the games [g0, g1): its own round counter, slots and network batch.
    const Tree v = whole ? e->t : group_view(e, g0, g1, 0);
    const size_t first = (size_t)g0 * e->t.inflight, n = (size_t)(g1 - g0) * e->t.inflight;
    float *feat = e->d_features + first * (1 + 2 * e->t.H) * 64;
    void *x0 = nullptr;
    if ((st = encode_target(e, n, 0, &x0))) { e->in_search = false; return st; }
    for (int r = 0; r < rounds && !st; ++r) {
        e->round = r;
        cudaEvent_t cb = nullptr, ce = nullptr;
        if (e->profiling) { cudaEventRecord(next_event(e, k), e->stream); cb = next_event(e, k + 2); ce = next_event(e, k + 3); }
        { NvtxRange r("oz: round select+encode"); launch_select(v, e->stream); launch_encode(v, feat, e->stream, false, x0); }
        if (e->profiling) cudaEventRecord(next_event(e, k + 1), e->stream);
        { NvtxRange r("oz: round network"); st = eval_slots(e, first, n, e->stream, 0, net_prof ? cb : nullptr, net_prof ? ce : nullptr, x0 != nullptr); }
        if (e->profiling) cudaEventRecord(next_event(e, k + 4), e->stream);
        if (!st) { NvtxRange r("oz: round expand+backup"); launch_expand(v, e->d_policy + first * 65, e->d_value + first, e->stream); }
        if (e->profiling) { cudaEventRecord(next_event(e, k + 5), e->stream); k += 6; }
    }
    if (st) { e->in_search = false; return st; }
    if (whole) e->warmed = true;
    st = end_search(e, v, active);
    if (!st && e->profiling) { e->prof_events = k; e->prof_pipelined = false; e->prof_net = net_prof; if (!e->defer_error_check) read_profile(e); }
    return st;
}

int oz_engine_search(oz_engine *e) { return search_games(e, 0, e->t.G); }

int oz_engine_search_range(oz_engine *e, int first_game, int num_games) {
    if (first_game < 0 || num_games < 1 || first_game + num_games > e->t.G)
        return fail(OZ_ERR_OUT_OF_RANGE, "Expected 0 <= first_game and first_game + num_games <= " + std::to_string(e->t.G) + ", but got " +
                                             std::to_string(first_game) + " + " + std::to_string(num_games) + ".");
    if (e->cfg.evaluator == OZ_EVAL_EXTERNAL) return fail(OZ_ERR_STATE, "oz_engine_search_range needs a built-in evaluator");
    return search_games(e, first_game, first_game + num_games);
}

int oz_engine_root_stats(oz_engine *e, int32_t *n_children, int32_t *counts, float *q, int32_t *root_n) {
    OZ_TRY(set_device(e));
    const int G = e->t.G;
    launch_root_stats(e->t, e->d_stat_n, e->d_stat_counts, e->d_stat_q, e->d_stat_rootn, e->stream);
    if (n_children) OZ_CUDA(cudaMemcpyAsync(n_children, e->d_stat_n, G * 4, cudaMemcpyDeviceToHost, e->stream));
    if (counts) OZ_CUDA(cudaMemcpyAsync(counts, e->d_stat_counts, (size_t)G * kStride * 4, cudaMemcpyDeviceToHost, e->stream));
    if (q) OZ_CUDA(cudaMemcpyAsync(q, e->d_stat_q, (size_t)G * kStride * 4, cudaMemcpyDeviceToHost, e->stream));
    if (root_n) OZ_CUDA(cudaMemcpyAsync(root_n, e->d_stat_rootn, G * 4, cudaMemcpyDeviceToHost, e->stream));
    OZ_CUDA(cudaStreamSynchronize(e->stream));
    return OZ_OK;
}

int oz_engine_self_play_data(oz_engine *e, int game, float *features, float *policy) {
    OZ_TRY(set_device(e));
    if (game < 0 || game >= e->t.G) return fail(OZ_ERR_OUT_OF_RANGE, "game index out of range");
    std::vector<oz_position> pos(e->t.G);
    OZ_TRY(oz_engine_positions(e, pos.data()));
    if (pos[game].player == 0)   // mcts.cpp:64-68
        return fail(OZ_ERR_INVALID_ARGUMENT, "Self-play data cannot be generated from a terminal position.");
    std::vector<int32_t> nch(e->t.G);
    OZ_TRY(oz_engine_root_stats(e, nch.data(), nullptr, nullptr, nullptr));
    if (nch[game] == 0)          // mcts.cpp:69-71
        return fail(OZ_ERR_INVALID_ARGUMENT, "The root node has not been expanded yet.");
    sync_params(e);
    launch_self_play_data(e->t, game, e->d_spd_f, e->d_spd_p, e->stream);
    const size_t nf = (size_t)8 * (1 + 2 * e->t.H) * 64;
    OZ_CUDA(cudaMemcpyAsync(features, e->d_spd_f, nf * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    OZ_CUDA(cudaMemcpyAsync(policy, e->d_spd_p, 8 * 65 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    OZ_CUDA(cudaStreamSynchronize(e->stream));
    return OZ_OK;
}

static int apply_device_actions(oz_engine *e) {
    NvtxRange nvtx_apply("oz: re-root (apply actions, harvest)");
    OZ_CUDA(cudaMemsetAsync(e->d_unfinished, 0, 2 * sizeof(int32_t), e->stream));
    launch_apply(e->t, e->d_actions, e->d_unfinished, e->stream);
    int32_t *unf2 = e->h_pin + 1;
    e->h_pin[3] = 0;
    OZ_CUDA(cudaMemcpyAsync(unf2, e->d_unfinished, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    if (e->done.capacity > 0) {            // rolling pool: completed games leave, their slots restart with unplayed games
        launch_harvest(e->t, e->d_records, e->max_plies, e->done, e->pool_total, e->stream);
        OZ_CUDA(cudaMemcpyAsync(e->h_pin + 3, e->done.counters + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    }
    OZ_CUDA(cudaStreamSynchronize(e->stream));         // the ply's one host synchronisation
    const int32_t unf = unf2[0], assigned = e->h_pin[3];
    e->tree_bound = std::max<long>(1, unf2[1]);       // exact: the largest live tree after the re-root
    e->ctr.games_finished += (uint64_t)std::max(0, e->active - unf);
    e->active = unf + assigned;
    return OZ_OK;
}

int oz_engine_apply_actions(oz_engine *e, const int32_t *actions) {
    OZ_TRY(set_device(e));
    if (e->in_search) return fail(OZ_ERR_STATE, "cannot apply actions during a search");
    const int G = e->t.G;
    std::vector<oz_position> pos(G);
    OZ_TRY(oz_engine_positions(e, pos.data()));
    for (int g = 0; g < G; ++g) {   // validation of mcts.cpp:115-138, same texts
        const int a = actions[g];
        if (a < 0 && G > 1) continue;
        if (!(0 <= a && a < 65)) return fail(OZ_ERR_OUT_OF_RANGE, "Expected 0 <= action < 65, but got " + std::to_string(a) + ".");
        if (a == 64) {
            if (pos[g].player == 0) return fail(OZ_ERR_INVALID_ARGUMENT, "Pass is not allowed in a terminal position.");
            if (pos[g].legal != 0) return fail(OZ_ERR_INVALID_ARGUMENT, "Pass is not allowed when there are legal moves.");
        } else if (((1ull << (63 - a)) & pos[g].legal) == 0) {
            return fail(OZ_ERR_INVALID_ARGUMENT, std::to_string(a) + " is not a legal action.");
        }
    }
    OZ_CUDA(cudaMemcpyAsync(e->d_actions, actions, G * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    OZ_TRY(apply_device_actions(e));
    e->ctr.plies += 1;
    return OZ_OK;
}

int oz_engine_choose_actions(oz_engine *e, int32_t *actions) {
    OZ_TRY(set_device(e));
    sync_params(e);
    launch_choose(e->t, e->d_actions, e->stream);
    OZ_CUDA(cudaMemcpyAsync(actions, e->d_actions, e->t.G * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    OZ_CUDA(cudaStreamSynchronize(e->stream));
    return OZ_OK;
}

int oz_engine_play_ply(oz_engine *e, int32_t *actions, int32_t *visit_counts, int32_t *unfinished) {
    NvtxRange nvtx_ply("oz: ply");
    e->defer_error_check = true;           // one host synchronisation per ply: at the end of apply_device_actions
    const int st = oz_engine_search(e);
    e->defer_error_check = false;
    OZ_TRY(st);
    const size_t G = (size_t)e->t.G;
    if (visit_counts) {
        launch_root_stats(e->t, e->d_stat_n, e->d_stat_counts, nullptr, nullptr, e->stream);
        OZ_CUDA(cudaMemcpyAsync(e->h_counts, e->d_stat_counts, G * kStride * 4, cudaMemcpyDeviceToHost, e->stream));
    }
    { NvtxRange r("oz: choose+record"); launch_choose(e->t, e->d_actions, e->stream); launch_record(e->t, e->d_actions, e->d_records, e->max_plies, e->stream); }
    if (actions) OZ_CUDA(cudaMemcpyAsync(e->h_actions, e->d_actions, G * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    OZ_TRY(apply_device_actions(e));
    read_profile(e);                       // the search's phase events have completed now
    OZ_TRY(device_error_status(e));        // the search's error bits arrived with the same synchronisation
    if (visit_counts) std::memcpy(visit_counts, e->h_counts, G * kStride * 4);
    if (actions) std::memcpy(actions, e->h_actions, G * sizeof(int32_t));
    e->ctr.plies += 1;
    if (unfinished) *unfinished = e->active;
    return OZ_OK;
}

int oz_engine_records(oz_engine *e, int32_t *plies, oz_ply_record *records, int max_plies) {
    OZ_TRY(set_device(e));
    const int G = e->t.G;
    std::vector<int32_t> p(G);
    OZ_CUDA(cudaMemcpy(p.data(), e->t.g_ply, G * 4, cudaMemcpyDeviceToHost));
    for (int g = 0; g < G; ++g) {
        if (p[g] > e->max_plies) p[g] = e->max_plies;
        if (plies) plies[g] = p[g];
        if (records) {
            const int n = std::min(p[g], max_plies);
            if (n > 0) OZ_CUDA(cudaMemcpy(records + (size_t)g * max_plies, e->d_records + (size_t)g * e->max_plies, (size_t)n * sizeof(oz_ply_record), cudaMemcpyDeviceToHost));
        }
    }
    return OZ_OK;
}

static int training_samples_from_store(oz_engine *e, float *features, float *policy, float *value, size_t capacity, size_t *n_samples) {
    int32_t rows = 0;
    OZ_CUDA(cudaMemcpy(&rows, e->done.counters, sizeof(rows), cudaMemcpyDeviceToHost));
    std::vector<int32_t> plies((size_t)std::max(rows, 1));
    if (rows > 0) OZ_CUDA(cudaMemcpy(plies.data(), e->done.plies, (size_t)rows * 4, cudaMemcpyDeviceToHost));
    std::vector<int64_t> base((size_t)std::max(rows, 1));
    size_t total = 0;
    for (int i = 0; i < rows; ++i) { base[i] = (int64_t)total; total += (size_t)plies[i] * 8; }
    *n_samples = total;
    if (!features || !policy || !value || total == 0) return OZ_OK;
    if (capacity < total) return fail(OZ_ERR_INVALID_ARGUMENT, "sample buffers too small");
    const size_t nf = (size_t)(1 + 2 * e->t.H) * 64;
    float *d_f = nullptr, *d_p = nullptr, *d_v = nullptr;
    int64_t *d_base = nullptr;
    int st = OZ_OK;
    do {
        if ((st = dalloc(&d_f, total * nf)) || (st = dalloc(&d_p, total * 65)) || (st = dalloc(&d_v, total)) || (st = dalloc(&d_base, rows))) break;
        cudaMemcpy(d_base, base.data(), (size_t)rows * 8, cudaMemcpyHostToDevice);
        launch_training_samples_rows(rows, e->t.H, e->done.records, e->done.plies, d_base, e->done.p1, e->done.p2, e->max_plies, d_f, d_p, d_v, e->stream);
        cudaError_t err = cudaStreamSynchronize(e->stream);
        if (err != cudaSuccess) { st = fail(OZ_ERR_CUDA, std::string("training_samples: ") + cudaGetErrorString(err)); break; }
        cudaMemcpy(features, d_f, total * nf * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(policy, d_p, total * 65 * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(value, d_v, total * sizeof(float), cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(d_f); cudaFree(d_p); cudaFree(d_v); cudaFree(d_base);
    return st;
}

int oz_engine_training_samples(oz_engine *e, float *features, float *policy, float *value, size_t capacity, size_t *n_samples) {
    OZ_TRY(set_device(e));
    if (e->done.capacity > 0) return training_samples_from_store(e, features, policy, value, capacity, n_samples);
    const int G = e->t.G;
    std::vector<int32_t> plies(G), status(G);
    OZ_CUDA(cudaMemcpy(plies.data(), e->t.g_ply, G * 4, cudaMemcpyDeviceToHost));
    OZ_CUDA(cudaMemcpy(status.data(), e->t.g_status, G * 4, cudaMemcpyDeviceToHost));
    std::vector<int64_t> base(G, -1);
    size_t total = 0;
    for (int g = 0; g < G; ++g) if (status[g] & 1) { base[g] = (int64_t)total; total += (size_t)std::min(plies[g], e->max_plies) * 8; }
    *n_samples = total;
    if (!features || !policy || !value || total == 0) return OZ_OK;
    if (capacity < total) return fail(OZ_ERR_INVALID_ARGUMENT, "sample buffers too small");
    const size_t nf = (size_t)(1 + 2 * e->t.H) * 64;
    float *d_f = nullptr, *d_p = nullptr, *d_v = nullptr;
    int64_t *d_base = nullptr;
    int32_t *d_plies = nullptr;
    int st = OZ_OK;
    do {
        if ((st = dalloc(&d_f, total * nf)) || (st = dalloc(&d_p, total * 65)) || (st = dalloc(&d_v, total)) || (st = dalloc(&d_base, G)) || (st = dalloc(&d_plies, G))) break;
        for (int g = 0; g < G; ++g) plies[g] = std::min(plies[g], e->max_plies);
        cudaMemcpy(d_base, base.data(), G * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(d_plies, plies.data(), G * 4, cudaMemcpyHostToDevice);
        launch_training_samples(e->t, e->d_records, d_plies, d_base, e->max_plies, d_f, d_p, d_v, e->stream);
        cudaError_t err = cudaStreamSynchronize(e->stream);
        if (err != cudaSuccess) { st = fail(OZ_ERR_CUDA, std::string("training_samples: ") + cudaGetErrorString(err)); break; }
        cudaMemcpy(features, d_f, total * nf * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(policy, d_p, total * 65 * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(value, d_v, total * sizeof(float), cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(d_f); cudaFree(d_p); cudaFree(d_v); cudaFree(d_base); cudaFree(d_plies);
    return st;
}

// The same expansion for records that travelled (e.g. gathered from other ranks as 176-byte rows instead of 4 612-byte
// samples): the final discs of a game follow from its last record and action, so the rows are self-contained.
int oz_training_samples_from_records(int device, const oz_ply_record *records, const int32_t *plies, int n_games, int max_plies,
                                     int history_size, float *features, float *policy, float *value, size_t capacity, size_t *n_samples) {
    if (n_games < 0 || max_plies < 1) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected n_games >= 0 and max_plies >= 1.");
    if (history_size < 1 || history_size > kMaxHist) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected 1 <= history_size <= " + std::to_string(kMaxHist) + ".");
    std::vector<int64_t> base((size_t)std::max(n_games, 1));
    std::vector<int32_t> pl((size_t)std::max(n_games, 1));
    std::vector<uint64_t> f1((size_t)std::max(n_games, 1)), f2((size_t)std::max(n_games, 1));
    size_t total = 0;
    for (int g = 0; g < n_games; ++g) {
        pl[g] = std::max(0, std::min(plies[g], max_plies));
        base[g] = (int64_t)total;
        total += (size_t)pl[g] * 8;
        f1[g] = kInitialP1; f2[g] = kInitialP2;
        if (pl[g] > 0) {                                   // final position = last recorded position + its action (train.py:438-439)
            const oz_ply_record &r = records[(size_t)g * max_plies + pl[g] - 1];
            Position p{r.player, r.p1, r.p2, 0, 0};
            const uint64_t me = r.player == 1 ? r.p1 : r.p2, opp = r.player == 1 ? r.p2 : r.p1;
            p.legal = legal_moves(me, opp);
            if (p.legal == 0) p.next_legal = legal_moves(opp, me);
            const Position q = apply_action(p, r.action);
            f1[g] = q.p1; f2[g] = q.p2;
        }
    }
    *n_samples = total;
    if (!features || !policy || !value || total == 0) return OZ_OK;
    if (capacity < total) return fail(OZ_ERR_INVALID_ARGUMENT, "sample buffers too small");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(OZ_ERR_CUDA, "no CUDA device: the engine has no CPU path");
    OZ_CUDA(cudaSetDevice(device));
    const size_t nf = (size_t)(1 + 2 * history_size) * 64, nrec = (size_t)n_games * max_plies;
    float *d_f = nullptr, *d_p = nullptr, *d_v = nullptr;
    int64_t *d_base = nullptr;
    int32_t *d_pl = nullptr;
    uint64_t *d_f1 = nullptr, *d_f2 = nullptr;
    PlyRecordDev *d_rec = nullptr;
    int st = OZ_OK;
    do {
        if ((st = dalloc(&d_f, total * nf)) || (st = dalloc(&d_p, total * 65)) || (st = dalloc(&d_v, total)) || (st = dalloc(&d_base, n_games)) ||
            (st = dalloc(&d_pl, n_games)) || (st = dalloc(&d_f1, n_games)) || (st = dalloc(&d_f2, n_games)) || (st = dalloc(&d_rec, nrec))) break;
        cudaMemcpy(d_rec, records, nrec * sizeof(PlyRecordDev), cudaMemcpyHostToDevice);
        cudaMemcpy(d_base, base.data(), (size_t)n_games * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(d_pl, pl.data(), (size_t)n_games * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d_f1, f1.data(), (size_t)n_games * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(d_f2, f2.data(), (size_t)n_games * 8, cudaMemcpyHostToDevice);
        launch_training_samples_rows(n_games, history_size, d_rec, d_pl, d_base, d_f1, d_f2, max_plies, d_f, d_p, d_v, nullptr);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { st = fail(OZ_ERR_CUDA, std::string("training_samples_from_records: ") + cudaGetErrorString(err)); break; }
        cudaMemcpy(features, d_f, total * nf * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(policy, d_p, total * 65 * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(value, d_v, total * sizeof(float), cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(d_f); cudaFree(d_p); cudaFree(d_v); cudaFree(d_base); cudaFree(d_pl); cudaFree(d_f1); cudaFree(d_f2); cudaFree(d_rec);
    return st;
}

int oz_engine_counters(oz_engine *e, oz_counters *out) {
    OZ_TRY(set_device(e));
    unsigned long long sc[4 * kStatCopies], s[4] = {0, 0, 0, 0};
    OZ_CUDA(cudaMemcpy(sc, e->t.stats, sizeof(sc), cudaMemcpyDeviceToHost));
    for (int k = 0; k < kStatCopies; ++k) for (int i = 0; i < 4; ++i) s[i] += sc[4 * k + i];
    e->ctr.nodes_expanded = s[0]; e->ctr.path_nodes = s[1]; e->ctr.children_read = s[2];
    e->ctr.kernel_launches = g_kernel_launches;
    *out = e->ctr;
    return OZ_OK;
}

int oz_engine_set_profiling(oz_engine *e, int on) { e->profiling = on != 0; return OZ_OK; }
int oz_engine_set_pipeline(oz_engine *e, int groups, int pdl) {
    if (groups < 0 || groups > kMaxGroups) return fail(OZ_ERR_INVALID_ARGUMENT, "Expected 0 <= groups <= " + std::to_string(kMaxGroups) + ", but got " + std::to_string(groups) + ".");
    if (e->in_search) return fail(OZ_ERR_STATE, "cannot change the pipeline during a search");
    e->pipeline = groups; e->pipeline_pdl = pdl != 0;
    return OZ_OK;
}
int oz_engine_set_graphs(oz_engine *e, int mode) { e->graph_mode = mode; if (mode == 0) drop_graph(e); return OZ_OK; }
void *oz_engine_stream(oz_engine *e) { return (void *)e->stream; }

}  // extern "C"
