// Device-resident search trees for lockstep games: struct-of-arrays node pools and the
// kernels that walk them.  reference: cpp/src/lib/search_thread.cpp, cpp/src/lib/mcts.cpp,
// cpp/src/include/search_node.h (AoS nodes + unique_ptr children there).
//
// Layout in HBM.  Every game owns two arenas of `cap` nodes (the live one and the target
// of the next re-root compaction).  Node id = index inside the live arena; global slot =
// (game*2 + which)*cap + id.  A node stores the edge that leads INTO it (search_node.h:31-45):
//   n_p1,n_p2   u64   discs (absolute colours)
//   n_meta      i32   player (bits 0-1: 0 terminal) | child count << 8
//   n_parent    i32   -1 for the root
//   n_child     i32   id of the first child, -1 if childless.  Children are created
//                     together at expansion, so they are CONTIGUOUS and in ascending action
//                     order: a warp reads a node's child stats with one coalesced access
//                     per array.
//   n_N i32, n_W f32, n_Q f32, n_P f32   visit count, total / mean action value, prior
// (player, p1, p2) determine legal / next_legal (position.h:349-358), so they are recomputed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oz {

constexpr int kMaxDepth = 128;   // longest root-to-leaf path kept per slot
constexpr int kMaxHist = 16;     // retired ancestors kept per game (history_size <= 16)
constexpr int kStride = 36;      // OZ_MAX_CHILDREN
constexpr int kStatCopies = 32;  // copies of the engine counters (Tree::stats)

enum TreeError { kErrArenaFull = 1, kErrRateTable = 2, kErrDepth = 4 };

struct Tree {
    // node pools
    uint64_t *n_p1, *n_p2;
    int32_t *n_meta, *n_parent, *n_child, *n_N;
    float *n_W, *n_Q, *n_P;
    // per game
    int32_t *g_root, *g_which, *g_count, *g_status, *g_ply;
    uint32_t *g_search;                  // searches since reset (RNG counter)
    uint32_t *g_gameid;                  // global game id played in this slot (RNG streams are keyed by it)
    uint64_t *h_p1, *h_p2;               // [G][kMaxHist] retired ancestors, most recent first
    int32_t *h_player, *h_n;             // [G][kMaxHist], [G]
    uint64_t *r_p1, *r_p2; int32_t *r_player;   // root position mirror [G]
    // per slot (game * inflight + i)
    int32_t *s_leaf, *s_tf, *s_len, *s_path;    // s_path: [slot][kMaxDepth]
    // misc
    const float *rate;                   // exploration rate by node visit count
    int rate_len;
    const float *gamma_table; int gamma_len;    // injected Dirichlet draws (tests)
    int32_t *round;                      // round of the current search (device-side, so one CUDA graph serves every round)
    int32_t *error;                      // TreeError bits
    // [kStatCopies][4]: [0] nodes expanded, [1] path nodes walked by select, [2] child statistics read by select.  A warp adds
    // to copy (its game % kStatCopies): a thousand games adding to one address serialise in L2 at the end of every select /
    // expand launch; the host sums the copies.
    unsigned long long *stats;
    // configuration
    int G, cap, H, inflight;
    float eps, alpha;
    uint32_t seed_lo, seed_hi, first_game_id;
    int sym_mode, temperature_plies;
    float inv_temperature;               // 1 / self_play_temperature (train.py:423)
};

struct PlyRecordDev {     // matches oz_ply_record
    uint64_t p1, p2;
    int32_t player, action, n_children;
    int32_t visit_counts[kStride];
};

void launch_reset(const Tree &t, cudaStream_t s);
// select / encode read the round from *t.round; expand increments it.
// narrow = small blocks (one game per block) that co-reside with a convolution CTA: the pipelined search.
// pdl = launch with programmatic dependent launch (launch_kernel, oz_common.cuh): the launch-bound searches chain the
// kernels of all their rounds this way, so a kernel's launch latency and set-up overlap its predecessor's execution.
void launch_select(const Tree &t, cudaStream_t s, bool narrow = false, bool pdl = false);
// nhwc != nullptr: write bf16 NHWC [slot][64][32] (the native network's input tensor) instead of the float planes.
void launch_encode(const Tree &t, float *features, cudaStream_t s, bool narrow = false, void *nhwc = nullptr, bool pdl = false);
void launch_expand(const Tree &t, const float *policy, const float *value, cudaStream_t s, bool narrow = false, bool pdl = false);
void launch_end_search(const Tree &t, cudaStream_t s);
void launch_root_stats(const Tree &t, int32_t *n_children, int32_t *counts, float *q, int32_t *root_n, cudaStream_t s);
void launch_choose(const Tree &t, int32_t *actions, cudaStream_t s);
void launch_record(const Tree &t, const int32_t *actions, PlyRecordDev *records, int max_plies, cudaStream_t s);
// unfinished[0] += games still running, unfinished[1] = max(unfinished[1], nodes in the largest re-rooted tree).
void launch_apply(const Tree &t, const int32_t *actions, int32_t *unfinished, cudaStream_t s);
// *out = max(*out, largest g_count).
void launch_max_count(const Tree &t, int32_t *out, cudaStream_t s);
void launch_self_play_data(const Tree &t, int game, float *features, float *policy, cudaStream_t s);
void launch_eval_synth(const float *features, int planes, size_t n, float *policy, float *value, cudaStream_t s);
void launch_eval_hash(const float *features, int planes, size_t n, float *policy, float *value, cudaStream_t s);
// Training tensors from records (train.py:432-450).  sample_base[g] = first sample row of game g.
void launch_training_samples(const Tree &t, const PlyRecordDev *records, const int32_t *plies, const int64_t *sample_base,
                             int max_plies, float *features, float *policy, float *value, cudaStream_t s);
// Same for `rows` record rows with explicit final positions (store of completed games).
void launch_training_samples_rows(int rows, int H, const PlyRecordDev *records, const int32_t *plies, const int64_t *sample_base,
                                  const uint64_t *fin_p1, const uint64_t *fin_p2, int max_plies, float *features, float *policy,
                                  float *value, cudaStream_t s);
// Rolling game pool: finished games are moved to a store of completed games and their slots restart with the next
// unplayed game id.  `done_*` arrays have one row per completed game, in completion order (slot order within a ply).
struct DoneStore {
    PlyRecordDev *records;    // [capacity][max_plies]
    int32_t *plies;           // [capacity]
    uint32_t *gameid;         // [capacity]
    uint64_t *p1, *p2;        // [capacity] final discs (train.py:438-439: the outcome)
    int32_t *counters;        // [0] completed games, [1] next unplayed game id offset, [2] games assigned by the last harvest
    int32_t *slot_done, *slot_new;   // [G] scratch: done row / new game id offset of a slot (-1 none)
    int capacity;             // games in the pool (completed rows available)
};
void launch_harvest(const Tree &t, const PlyRecordDev *records, int max_plies, const DoneStore &d, int pool_total, cudaStream_t s);
// Move every game's live arena into pools with a different capacity (pool growth).
void launch_copy_arenas(const Tree &from, const Tree &to, cudaStream_t s);

}  // namespace oz
