/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's self-play hot path.
 *
 * Nothing under oracle/ is part of the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and only
 * as the checker (or the reported CPU baseline), never as the thing shipped.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against the unmodified reference compiled from /root/reference (oracle/_ref, built by
 * oracle/build_oracle.py) and against the golden vectors committed in tests/golden/
 * (generated from the reference by tests/golden/generate_golden.py).  The Dirichlet /
 * symmetry random streams of the reference cannot be seeded (search_thread.cpp:22), so
 * the noisy regime is pinned through the deterministic regime plus injected streams.
 *
 * Each function cites the reference file:line it restates (paths relative to the
 * reference root).  Compile with -ffp-contract=off: the reference build has no FMA
 * contraction and the PUCT comparisons depend on it.
 */
#ifndef OZ_ORACLE_H
#define OZ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- rules: cpp/src/include/position.h ---- */
typedef struct {
    int32_t  player;      /* 1 Black, 2 White, 0 terminal            position.h:144 */
    uint64_t p1, p2;      /* Black / White discs                      position.h:145-146 */
    uint64_t legal;       /* position.h:147 */
    uint64_t next_legal;  /* position.h:148 */
} ozo_position;

uint64_t ozo_get_legal_moves(uint64_t player_discs, uint64_t opponent_discs);
uint64_t ozo_get_flips(uint64_t move_mask, uint64_t player_discs, uint64_t opponent_discs);
void     ozo_initial_position(ozo_position *out);
void     ozo_apply_move(const ozo_position *p, uint64_t move_mask, ozo_position *out);
void     ozo_apply_pass(const ozo_position *p, ozo_position *out);
void     ozo_apply_action(const ozo_position *p, int action, ozo_position *out);
int      ozo_legal_actions(const ozo_position *p, int *actions /* >= 33 */);
uint64_t ozo_perft(const ozo_position *p, int depth);

/* ---- encoding: cpp/src/include/transformation.h ---- */
int  ozo_transform_action(int action, int transformation);
/* history[0] is the current position, history[k] the position k plies back; n_hist may
 * be smaller than history_size (missing planes are zero).  features: (1+2H)*64 floats. */
void ozo_positions_to_features(const ozo_position *history, int n_hist, float *features,
                               int history_size, int transformation);

/* ---- counter-based RNG shared with the device engine (engine spec, not reference) ---- */
void ozo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ---- search: cpp/src/lib/search_thread.cpp, cpp/src/lib/mcts.cpp ---- */
typedef void (*ozo_eval_fn)(void *ctx, const float *features, int batch, int planes,
                            float *policy /* batch*65 */, float *value /* batch */);

enum { OZO_EVAL_CALLBACK = 0, OZO_EVAL_SYNTH = 1, OZO_EVAL_HASH = 2 };
enum { OZO_SYM_PHILOX = 0, OZO_SYM_IDENTITY = 1 };

typedef struct ozo_mcts ozo_mcts;

ozo_mcts *ozo_mcts_create(int history_size, int num_simulations, int num_threads,
                          int batch_size, float c_puct_base, float c_puct_init,
                          float dirichlet_epsilon, float dirichlet_alpha);
void ozo_mcts_destroy(ozo_mcts *m);
/* RNG stream identity: (seed, game_id); gamma_table != NULL injects the Dirichlet draws
 * (value = table[philox_word % len]); otherwise Marsaglia-Tsang on Philox uniforms. */
void ozo_mcts_set_rng(ozo_mcts *m, uint64_t seed, uint32_t game_id, const float *gamma_table,
                      int gamma_table_len, int sym_mode);
void ozo_mcts_set_evaluator(ozo_mcts *m, int kind, ozo_eval_fn fn, void *ctx);
void ozo_mcts_reset_position(ozo_mcts *m);
void ozo_mcts_position(const ozo_mcts *m, ozo_position *out);
void ozo_mcts_search(ozo_mcts *m);
int  ozo_mcts_visit_counts(const ozo_mcts *m, int *counts /* >= 33 */);
int  ozo_mcts_mean_action_values(const ozo_mcts *m, float *values /* >= 33 */);
int  ozo_mcts_root_visit_count(const ozo_mcts *m);
/* 8 symmetries: features 8*(1+2H)*64 floats, policy 8*65 floats.  Returns 0, or -1 for a
 * terminal root, -2 for an unexpanded root (mcts.cpp:64-71). */
int  ozo_mcts_self_play_data(const ozo_mcts *m, float *features, float *policy);
/* Returns 0, -1 out of range, -2 illegal (mcts.cpp:115-138). */
int  ozo_mcts_apply_action(ozo_mcts *m, int action);
/* Engine action rule (stands in for train.py:421-430 with a counter-based stream):
 * ply < temperature_plies: sample proportional to visit counts (tau = 1); otherwise a
 * uniformly random argmax.  Returns the action. */
int  ozo_mcts_choose_action(const ozo_mcts *m, int ply, int temperature_plies);
long ozo_mcts_num_nodes(const ozo_mcts *m);

/* Built-in exact evaluators (also implemented as device kernels by the engine). */
void ozo_eval_synth(const float *features, int batch, int planes, float *policy, float *value);
void ozo_eval_hash(const float *features, int batch, int planes, float *policy, float *value);

#ifdef __cplusplus
}
#endif
#endif
